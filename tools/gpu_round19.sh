#!/bin/bash
# Rewritten HBM kernels (row-in-register softmax / xent, FSMN float4 windows, splice, BN single stats pass), SFU math in the
# recurrence / CTC, forwarder CLIs: full GPU test suite, per-class HBM bench, bench line, ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 1200 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 40 gpurun_out/t_$name.log | cut -c1-300 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run tests python -m pytest tests -q -m gpu -p no:cacheprovider
run hbm python tools/kernel_bench.py
cp gpurun_out/t_hbm.log gpurun_out/hbm_kernels.jsonl
run probe python tools/perf_probe.py recur
run ctcprobe python tools/perf_probe.py ctc
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'act_|softmax|xent|bn_|splice|fsmn|axpby|col_reduce' -c 60 -f -o gpurun_out/prof_hbm python tools/kernel_bench.py --once "" > gpurun_out/t_ncu_hbm.log 2>&1
echo "ncu_hbm exit=$?" >> gpurun_out/summary.txt
grep -v abnormal gpurun_out/summary.txt
