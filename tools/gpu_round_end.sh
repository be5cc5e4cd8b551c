#!/bin/bash
# One call for the round-end evidence on one B200: the full GPU test suite and smoke(), then everything profiles/ keeps
# (tools/gpu_profiles.sh: bench line, reference arm, HBM table, launch list, ncu summaries), then two small probes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/t_tests.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/t_smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/t_smoke.log
bash tools/gpu_profiles.sh
timeout -s KILL 300 python tools/perf_probe.py timing > gpurun_out/lstm_phase_timing.txt 2>&1
timeout -s KILL 300 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1
ASLP_LSTM_BWD_T=0 timeout -s KILL 300 python tools/config_bench.py > gpurun_out/config_bench_gatherall.log 2>&1
grep '^{' gpurun_out/config_bench.log gpurun_out/config_bench_gatherall.log | cut -c1-220
