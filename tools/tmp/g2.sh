cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash tools/gpu_ab_fwd.sh pr1 pr2 pr3s pr3f
timeout -s KILL 300 python tools/perf_probe.py timing > gpurun_out/lstm_phase_timing_pipe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_nnet_golden.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pipe_tests.log
