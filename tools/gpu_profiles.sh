#!/bin/bash
# Captures what profiles/ keeps (one B200): the bench line and the reference arm, the per-kernel HBM table, the launch list of
# the bench command, and ncu --set full summaries of the recurrence, GEMM, CTC and HBM kernels.  The big .ncu-rep files are
# summarised ON the box and removed (gpurun brings back at most 64 MiB); only the recurrence report travels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.log 2>&1; echo "bench exit=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref exit=$?"
python tools/kernel_bench.py > gpurun_out/hbm_kernels.jsonl 2>&1; echo "hbm exit=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/t_ncu_launches.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -f -o gpurun_out/prof_lstm_mma python tools/perf_probe.py lstm0 > gpurun_out/t_ncu_lstm.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_lstm_mma.ncu-rep > gpurun_out/lstm_mma_ncu_summary.txt
timeout -s KILL 600 ncu --set full --clock-control none -k regex:gemm_tf32 -c 2 -f -o gpurun_out/prof_gemm python tools/perf_probe.py gemm1 > gpurun_out/t_ncu_gemm.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_gemm.ncu-rep > gpurun_out/gemm_ncu_summary.txt; rm -f gpurun_out/prof_gemm.ncu-rep
timeout -s KILL 600 ncu --set full --clock-control none -k regex:ctc_ -c 4 -f -o gpurun_out/prof_ctc python tools/perf_probe.py ctc1 > gpurun_out/t_ncu_ctc.log 2>&1
ASLP_CTC_SWEEP=block timeout -s KILL 600 ncu --set full --clock-control none -k regex:ctc_ -c 4 -f -o gpurun_out/prof_ctc4 python tools/perf_probe.py ctc1 > gpurun_out/t_ncu_ctc4.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_ctc4.ncu-rep > gpurun_out/ctc_four_launch_ncu_summary.txt; rm -f gpurun_out/prof_ctc4.ncu-rep
python tools/ncu_summary.py gpurun_out/prof_ctc.ncu-rep > gpurun_out/ctc_ncu_summary.txt; rm -f gpurun_out/prof_ctc.ncu-rep
timeout -s KILL 600 ncu --set full --clock-control none -k regex:'act_|softmax_reg|xent_reg|bn_|splice|fsmn|axpby|col_reduce' -c 60 -f -o gpurun_out/prof_hbm python tools/kernel_bench.py --once "" > gpurun_out/t_ncu_hbm.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_hbm.ncu-rep > gpurun_out/hbm_ncu_summary.txt; rm -f gpurun_out/prof_hbm.ncu-rep
# the CNN front end's kernels (rewritten late in round 2): their own ncu summary
timeout -s KILL 600 ncu --set full --clock-control none -k regex:'conv_|maxpool_' -c 8 -f -o gpurun_out/prof_conv python tools/kernel_bench.py --once conv > gpurun_out/t_ncu_conv.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:'conv_|maxpool_' -c 8 -f -o gpurun_out/prof_pool python tools/kernel_bench.py --once maxpool > gpurun_out/t_ncu_pool.log 2>&1
(python tools/ncu_summary.py gpurun_out/prof_conv.ncu-rep; python tools/ncu_summary.py gpurun_out/prof_pool.ncu-rep) > gpurun_out/conv_pool_ncu_summary.txt; rm -f gpurun_out/prof_conv.ncu-rep gpurun_out/prof_pool.ncu-rep
# launch lists of one minibatch of the launch-bound configurations (enqueued, not replayed, so that every kernel is listed)
ASLP_STEP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/launches_cfg1.csv python tools/config_bench.py cfg1 > gpurun_out/t_ncu_cfg1.log 2>&1
ASLP_STEP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 5000 --csv --log-file gpurun_out/launches_cfg4.csv python tools/config_bench.py cfg4 > gpurun_out/t_ncu_cfg4.log 2>&1
python tools/config_bench.py 2>&1 | grep '^{' > gpurun_out/config_bench.jsonl
ASLP_STEP_GRAPH=0 python tools/config_bench.py 2>&1 | grep '^{' > gpurun_out/config_bench_enqueued.jsonl
ASLP_STEP_GRAPH=0 ASLP_FUSE_EPILOGUE=0 python tools/config_bench.py 2>&1 | grep '^{' > gpurun_out/config_bench_enqueued_unfused.jsonl
tail -1 gpurun_out/bench_ours.log | cut -c1-250; tail -1 gpurun_out/bench_ref.log | cut -c1-300
nproc > gpurun_out/nproc.txt; du -sh gpurun_out
