#!/bin/bash
# Round-end check on one B200: full GPU test suite, smoke, the bench line (with the cpu_baseline leg) and the reference arm.
# (profiles: tools/gpu_profiles.sh; scaling: tools/gpu_scale8.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 1500 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 8 gpurun_out/t_$name.log | cut -c1-300 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
run tests python -m pytest tests -q -m gpu -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.log 2>&1; echo "bench exit=$?" >> gpurun_out/summary.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref exit=$?" >> gpurun_out/summary.txt
grep -v abnormal gpurun_out/summary.txt
tail -1 gpurun_out/bench_ours.log | cut -c1-250; tail -1 gpurun_out/bench_ref.log | cut -c1-200
