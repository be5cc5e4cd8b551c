#!/bin/bash
# GEMM with the coalesced (staged) epilogue: correctness + timing; per-class HBM kernel bench + ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 900 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 40 gpurun_out/t_$name.log | cut -c1-300 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run gemm python -m pytest tests/test_gpu_gemm.py -q -m gpu -x -p no:cacheprovider
run golden python -m pytest tests/test_gpu_nnet_golden.py -q -m gpu -x -p no:cacheprovider
run probe python tools/perf_probe.py
run hbm python tools/kernel_bench.py
cp gpurun_out/t_hbm.log gpurun_out/hbm_kernels.jsonl
run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:'act_|softmax|xent|bn_|splice|fsmn|axpby|col_reduce' -c 40 -f -o gpurun_out/prof_hbm python tools/kernel_bench.py --once "" > gpurun_out/t_ncu_hbm.log 2>&1
echo "ncu_hbm exit=$?" >> gpurun_out/summary.txt
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 2 -f -o gpurun_out/prof_gemm2 python tools/perf_probe.py gemm1 > gpurun_out/t_ncu_gemm2.log 2>&1
cat gpurun_out/summary.txt | grep -v abnormal
