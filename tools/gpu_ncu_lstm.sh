#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -f -o gpurun_out/prof_lstm_mma python tools/perf_probe.py lstm0 > gpurun_out/t_ncu_lstm.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/t_ncu_lstm.log
