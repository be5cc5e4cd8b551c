#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:lstm_bwd -c 1 -f -o gpurun_out/prof_lstm_bwd2 python tools/perf_probe.py lstm0 > gpurun_out/t_ncu_lstm2.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/t_ncu_lstm2.log
