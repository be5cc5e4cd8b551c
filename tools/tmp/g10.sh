cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc.py tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/t_tests.log
ASLP_CTC_SWEEP=fused timeout 300 python tools/perf_probe.py ctc 2>&1 | grep '^{'
