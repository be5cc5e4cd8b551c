#!/bin/bash
# First-contact run on the B200 box: every test group in its own process under `timeout`,
# so a trapped kernel (dead CUDA context) or a hang cannot take the other groups with it.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export ASLP_B200_ALLOW_MISSING=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
run() { # name, pytest args...
  local name=$1; shift
  timeout -s KILL 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/t_$name.log 2>&1
  echo "$name exit=$?" >> gpurun_out/summary.txt
  tail -n 4 gpurun_out/t_$name.log | sed "s/^/[$name] /" >> gpurun_out/summary.txt
}
rm -f gpurun_out/summary.txt
run pointwise tests/test_gpu_pointwise.py
run ctc tests/test_gpu_ctc.py
run lstm tests/test_gpu_lstm.py
run nnet_golden tests/test_gpu_nnet_golden.py
run gemm_fp32 tests/test_gpu_gemm.py -k "all_layouts and fp32"
run gemm_nt_tf32 tests/test_gpu_gemm.py -k "all_layouts and NT and -tf32"
run gemm_nt_3x tests/test_gpu_gemm.py -k "all_layouts and NT and x3tf32"
run gemm_nn tests/test_gpu_gemm.py -k "all_layouts and NN and tf32"
run gemm_tn tests/test_gpu_gemm.py -k "all_layouts and TN and tf32"
run gemm_tt tests/test_gpu_gemm.py -k "all_layouts and TT and tf32"
run gemm_rest tests/test_gpu_gemm.py -k "not all_layouts"
timeout -s KILL 600 python tools/perf_probe.py > gpurun_out/perf_probe.jsonl 2> gpurun_out/perf_probe.err
echo "perf_probe exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/perf_probe.jsonl
