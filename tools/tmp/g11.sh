cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash tools/gpu_profiles.sh
timeout -s KILL 300 python tools/perf_probe.py timing > gpurun_out/lstm_phase_timing.txt 2>&1
ls -la gpurun_out | head -50
