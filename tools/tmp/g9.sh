cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/t_tests.log
python tools/config_bench.py 2>&1 | grep '^{' | tee gpurun_out/config_bench6.log
ASLP_ASYNC_WGRAD=0 python tools/config_bench.py 2>&1 | grep '^{' | tee gpurun_out/config_bench6_noside.log
