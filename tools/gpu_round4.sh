#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 900 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 6 gpurun_out/t_$name.log | cut -c1-400 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run golden python -m pytest tests/test_gpu_nnet_golden.py -q -m gpu -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --steps 5 --warmup 3
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run ncu_lstm ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -o gpurun_out/prof_lstm python tools/perf_probe.py lstm
cat gpurun_out/summary.txt
