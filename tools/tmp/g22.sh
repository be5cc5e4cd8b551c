cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_cfg2.csv python tools/config_bench.py cfg2 > gpurun_out/t_ncu_cfg2.log 2>&1
python - <<'PY'
import csv, re, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_cfg2.csv')) if len(r)>12 and r[0].isdigit()]
names=[re.sub(r'\(.*','',r[4]).replace('void ','').replace('<unnamed>::','') for r in rows]; durs=[float(r[-1].replace(',','')) for r in rows]; grid=[r[8] for r in rows]
per=len(rows)//32
last=list(zip(names,durs,grid))[-per:]
print('launches per minibatch', per, 'sum %.1f us'%(sum(d for _,d,_ in last)/1e3))
agg=collections.OrderedDict()
for k,d,g in last:
    a=agg.setdefault(k+' '+g,[0,0.0]); a[0]+=1; a[1]+=d
for k,(c,d) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]: print('   %-70s x%-3d %8.1f us  avg %.1f'%(k[:70],c,d/1e3,d/1e3/c))
PY
