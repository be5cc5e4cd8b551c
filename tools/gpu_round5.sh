#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 900 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 6 gpurun_out/t_$name.log | cut -c1-400 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run lstm python -m pytest tests/test_gpu_lstm.py -q -m gpu -p no:cacheprovider
run golden python -m pytest tests/test_gpu_nnet_golden.py -q -m gpu -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run probe python tools/perf_probe.py recur
run bench python bench.py --steps 5 --warmup 3
ASLP_LSTM_FOLD_PROJECTION=0 run bench_twostep python bench.py --steps 5 --warmup 3
run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
