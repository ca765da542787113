#!/bin/bash
mkdir -p gpurun_out
for t in 0 1 2 3; do echo "TUNE $t"; VAENAR_XROW_TUNE=$t timeout 120 python tools/xrow_phases.py > gpurun_out/xrow_phases_t$t.log 2>&1; grep -E "kernel span|ffn issued|c ready" gpurun_out/xrow_phases_t$t.log | head -3; done
VAENAR_SINGLE_CHAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_r2_single.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[]
with open('gpurun_out/launches_r2_single.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict()
for row in r:
    k=(row['Kernel Name'][:60], row['Grid Size'])
    v=float(row['Metric Value'])
    u=row['Metric Unit']
    if u in ('ns','nsecond'): v/=1e3
    elif u in ('ms','msecond'): v*=1e3
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print('total us', round(tot,1))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{k[0]:62s} {k[1]:16s} n={a[0]:4d} tot={a[1]:8.1f} avg={a[1]/a[0]:6.1f} share={a[1]/tot:.3f}")
PY
