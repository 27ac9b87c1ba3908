#!/bin/bash
# round 2, batch 12 (1 GPU): push kernel v2 (4x4 blocks, 16-byte shared accesses) + packed collect: parity, timing, ncu durations
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "virtual" --tb=short 2>&1 | tail -4
timeout 300 python scripts/r02_slice_tail.py 12 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:slice_ -c 24 --csv --log-file gpurun_out/r02_slice_launches.csv python scripts/r02_slice_tail.py 12 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_slice_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); idi=hdr.index('ID')
byid={}
for r in rows[1:]:
    byid.setdefault((r[idi], r[ki][:30]), {})[r[mi]]=r[vi]
for (i,kname),m in list(byid.items()):
    print(i, kname, {k.split('.')[0][-22:]:v for k,v in m.items()})
PY
