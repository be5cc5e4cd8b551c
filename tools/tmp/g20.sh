cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:ctc_ --csv --log-file gpurun_out/ctc2048.csv python tools/perf_probe.py ctc > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ctc2048.csv')) if len(r)>12 and r[0].isdigit()]
cur={}
for r in rows:
    key=(r[0], r[4].split('(')[0][-40:], r[8])
    cur.setdefault(key,{})[r[-3]]=r[-1]
for k,v in list(cur.items())[-12:]:
    print(k, v)
PY
