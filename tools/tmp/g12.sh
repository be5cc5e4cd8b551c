cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin_mains.py -q -m gpu -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/t_dropin.log
python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline 2>gpurun_out/bench_err.log | tail -1 | cut -c1-200; grep -c "abnormal" gpurun_out/bench_err.log
