cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ASLP_STEP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/launches_cfg1.csv python tools/config_bench.py cfg1 > gpurun_out/t_ncu_cfg1.log 2>&1
ASLP_STEP_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 5000 --csv --log-file gpurun_out/launches_cfg4.csv python tools/config_bench.py cfg4 > gpurun_out/t_ncu_cfg4.log 2>&1
