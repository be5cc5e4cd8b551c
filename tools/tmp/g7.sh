cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_tests.log
python tools/config_bench.py 2>&1 | grep '^{' | tee gpurun_out/config_bench4.log
ASLP_STEP_GRAPH=0 python tools/config_bench.py 2>&1 | grep '^{' | tee gpurun_out/config_bench4_eager.log
ASLP_STEP_GRAPH=0 ASLP_FUSE_EPILOGUE=0 python tools/config_bench.py 2>&1 | grep '^{' | tee gpurun_out/config_bench4_eager_unfused.log
