cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/ctc_tests.log
for d in 0 1 3 8 11; do echo "== dbg $d"; ASLP_CTC_DBG=$d ASLP_CTC_SWEEP=fused timeout 300 python tools/perf_probe.py ctc 2>&1 | grep '^{' ; done | tee gpurun_out/ctc_probe.log
