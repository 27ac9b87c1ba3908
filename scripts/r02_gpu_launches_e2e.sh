#!/bin/bash
# round 2: launch list of the count bench including the end-to-end leg with the hybrid upload
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench_count_e2e.csv python bench.py --workload count --steps 2 --warmup 3 > gpurun_out/ncu_e2e.log 2>&1
tail -1 gpurun_out/ncu_e2e.log | cut -c1-300
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_bench_count_e2e.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value'); u=h.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    name=r[k].split('(')[0][:50]
    t=float(r[v].replace(',',''))*(1e-3 if r[u]=='ns' else 1.0 if r[u] in ('us','usecond') else 1e3)
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
for n,(c,t) in agg.items(): print('%-52s launches %4d  total %9.1f us  mean %8.1f us' % (n,c,t,t/c))
PY
