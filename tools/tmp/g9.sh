cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/t_tests.log
python tools/config_bench.py 2>&1 | grep '^{' | tee gpurun_out/config_bench5.log
python tools/gemm_bench.py 2>&1 | tail -1 > gpurun_out/gemm_bench5.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.log 2>&1; tail -1 gpurun_out/bench5.log | cut -c1-400
