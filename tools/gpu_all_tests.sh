#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 1500 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 15 gpurun_out/t_$name.log | cut -c1-800 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run all python -m pytest tests -q -m gpu -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
cat gpurun_out/summary.txt
