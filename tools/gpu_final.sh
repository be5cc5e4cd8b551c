#!/bin/bash
# Round-end evidence on one B200: full GPU test suite, the bench line (with the cpu_baseline leg) and the reference arm, the
# per-kernel HBM table, the launch list of the bench command, ncu --set full of the recurrence, GEMM, CTC and HBM kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 1500 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 12 gpurun_out/t_$name.log | cut -c1-400 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run tests python -m pytest tests -q -m gpu -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.log 2>&1; echo "bench exit=$?" >> gpurun_out/summary.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref exit=$?" >> gpurun_out/summary.txt
run hbm python tools/kernel_bench.py
cp gpurun_out/t_hbm.log gpurun_out/hbm_kernels.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/t_ncu_launches.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -f -o gpurun_out/prof_lstm_mma python tools/perf_probe.py lstm0 > gpurun_out/t_ncu_lstm.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 2 -f -o gpurun_out/prof_gemm python tools/perf_probe.py gemm1 > gpurun_out/t_ncu_gemm.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:ctc_ -c 4 -f -o gpurun_out/prof_ctc python tools/perf_probe.py ctc1 > gpurun_out/t_ncu_ctc.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:'act_|softmax_reg|xent_reg|bn_|splice|fsmn|axpby|col_reduce' -c 60 -f -o gpurun_out/prof_hbm python tools/kernel_bench.py --once "" > gpurun_out/t_ncu_hbm.log 2>&1
grep -v abnormal gpurun_out/summary.txt
tail -1 gpurun_out/bench_ours.log | cut -c1-250; tail -1 gpurun_out/bench_ref.log | cut -c1-300
nproc
