#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 900 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 6 gpurun_out/t_$name.log | cut -c1-1500 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run timing python tools/perf_probe.py timing
run golden python -m pytest tests/test_gpu_nnet_golden.py tests/test_gpu_cli.py tests/test_gpu_workers.py -q -m gpu -p no:cacheprovider
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
ASLP_ASYNC_WGRAD=0 run bench_sync python bench.py --steps 5 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
