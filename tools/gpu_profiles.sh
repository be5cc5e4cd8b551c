#!/bin/bash
# Captures what profiles/ keeps: launch list of the bench command, ncu --set full of the recurrence kernels and of one
# large GEMM, and the bench lines themselves.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/t_ncu_launches.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -f -o gpurun_out/prof_lstm_mma python tools/perf_probe.py lstm0 > gpurun_out/t_ncu_lstm.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 2 -f -o gpurun_out/prof_gemm python tools/perf_probe.py gemm1 > gpurun_out/t_ncu_gemm.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:ctc_ -c 4 -f -o gpurun_out/prof_ctc python tools/perf_probe.py ctc1 > gpurun_out/t_ncu_ctc.log 2>&1
tail -1 gpurun_out/bench_ours.log | cut -c1-300; tail -1 gpurun_out/bench_ref.log | cut -c1-300
nproc
