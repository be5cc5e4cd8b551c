cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc.py tests/test_gpu_ctc_eesen.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/ctc_tests.log
for m in fused warp block; do echo "== $m"; ASLP_CTC_SWEEP=$m timeout 300 python tools/perf_probe.py ctc 2>&1 | grep '^{' ; done | tee gpurun_out/ctc_probe.log
