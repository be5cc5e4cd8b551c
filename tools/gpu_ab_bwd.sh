#!/bin/bash
# A/B of the transposed backward's contraction form (product = fp16 scaled split, variant bwdtf32 = 3xTF32) plus the
# LSTM / golden / full-size parity tests on the product library.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2; do
  for p in ab_probe.py ab_probe_wide.py; do
    python tools/$p "product"
    ASLP_B200_CUDA_LIB=$PWD/kaldi-aslp_b200/libaslp_b200_bwdtf32.so python tools/$p "bwdtf32"
  done
done 2>&1 | grep variant | tee gpurun_out/ab_bwd.jsonl
timeout 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_nnet_golden.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/ab_bwd_tests.log
python tools/config_bench.py 2>&1 | tail -8 | tee gpurun_out/config_bench.log
