#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 900 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 16 gpurun_out/t_$name.log | cut -c1-400 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run gemm timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -p no:cacheprovider -x
run probe timeout 300 python tools/perf_probe.py
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
