#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 1200 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 45 gpurun_out/t_$name.log | cut -c1-300 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run dbg python tools/debug_fwd_lc.py
run mma ./tools/micro/mma_rate
run tests python -m pytest tests/test_gpu_pointwise.py tests/test_gpu_nnet_golden.py tests/test_gpu_cli.py -q -m gpu -p no:cacheprovider
run hbm python tools/kernel_bench.py
cp gpurun_out/t_hbm.log gpurun_out/hbm_kernels.jsonl
grep -v abnormal gpurun_out/summary.txt
