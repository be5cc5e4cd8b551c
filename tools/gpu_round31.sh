#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_lstm.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -4
for i in 1 2; do
timeout -s KILL 300 python tools/ab_probe.py "transposed bwd"
ASLP_LSTM_BWD_T=0 timeout -s KILL 300 python tools/ab_probe.py "gather-all bwd"
ASLP_B200_CUDA_LIB=$PWD/kaldi-aslp_b200/libaslp_b200_d300.so timeout -s KILL 300 python tools/ab_probe.py "transposed + 300 ns"
ASLP_B200_CUDA_LIB=$PWD/kaldi-aslp_b200/libaslp_b200_d600.so timeout -s KILL 300 python tools/ab_probe.py "transposed + 600 ns"
done
