#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_nnet_golden.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py tests/test_gpu_workers.py -q -m gpu -p no:cacheprovider 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t.log 2>&1; grep '^{' gpurun_out/bench_t.log | tail -1 | cut -c1-260
