#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_lstm.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -12
timeout -s KILL 300 python tools/ab_probe.py "transposed bwd"
ASLP_LSTM_BWD_T=0 timeout -s KILL 300 python tools/ab_probe.py "gather-all bwd"
timeout -s KILL 300 python tools/ab_probe.py "transposed bwd"
