cd "${GRAFT_REPO_ROOT:-/root/repo}"
python - <<'PY'
import sys, json
sys.path.insert(0,'tools')
import perf_probe as PP
for T in (20, 100, 400):
    for bwd in (False, True):
        r=PP.probe_lstm(T, 100, 512, 0, 1, bwd)
        print(json.dumps({"T":T,"bwd":bwd,"ms":round(r["ms"],4),"us_per_step":round(r["us_per_step"],2)}), flush=True)
PY
python tools/config_bench.py cfg2 2>&1 | grep '^{'
