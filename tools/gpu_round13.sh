#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 1500 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 30 gpurun_out/t_$name.log | cut -c1-500 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run full python -m pytest tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider --durations=8
cat gpurun_out/summary.txt
