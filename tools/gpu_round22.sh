#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 1500 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 30 gpurun_out/t_$name.log | cut -c1-600 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run tests python -m pytest tests -q -m gpu -p no:cacheprovider -x
run probe python tools/perf_probe.py recur
run timing python tools/perf_probe.py timing
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run hbm python tools/kernel_bench.py
cp gpurun_out/t_hbm.log gpurun_out/hbm_kernels.jsonl
grep -v abnormal gpurun_out/summary.txt
