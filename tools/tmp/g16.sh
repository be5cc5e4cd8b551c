cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash tools/gpu_ab_fwd.sh nopre
timeout -s KILL 300 python tools/perf_probe.py timing 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['us_per_step'],3), d['cycles_per_step_mean'])"
