cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/t_tests.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/t_smoke.log 2>&1; echo "smoke exit=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.log 2>&1; echo "bench exit=$?"
tail -1 gpurun_out/bench_ours.log | cut -c1-3000
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/t_ncu_launches.log 2>&1
timeout -s KILL 300 python tools/perf_probe.py timing > gpurun_out/lstm_phase_timing.txt 2>&1
timeout -s KILL 300 python tools/config_bench.py > gpurun_out/config_bench.log 2>&1
grep '^{' gpurun_out/config_bench.log | cut -c1-300
cat gpurun_out/lstm_phase_timing.txt | cut -c1-400
