cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/t_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.log 2>&1; tail -1 gpurun_out/bench_ours.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['guard_rejected_utts'])
for k in ('cfg1_minibatch','cfg2_minibatch','cfg4_minibatch','ctc_16_utts','ctc_2048_utts'):
    v=d['secondary'][k]; print(k, v.get('frames_per_s') or round(v.get('utts_per_s')), v.get('ms_per_minibatch') or round(v.get('ms'),3))"
