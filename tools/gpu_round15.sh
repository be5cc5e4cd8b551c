#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 900 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 12 gpurun_out/t_$name.log | cut -c1-500 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run tests python -m pytest tests/test_gpu_lstm.py tests/test_gpu_nnet_golden.py tests/test_gpu_fullsize.py tests/test_gpu_cli.py -q -m gpu -p no:cacheprovider
run probe python tools/perf_probe.py recur
run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
