#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 600 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 30 gpurun_out/t_$name.log | cut -c1-600 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run contract ./tools/micro/contract_rate
cat gpurun_out/summary.txt
