#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 1200 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 45 gpurun_out/t_$name.log | cut -c1-400 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run dbg python tools/debug_fwd_lc.py
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
grep -v abnormal gpurun_out/summary.txt
