#!/bin/bash
# round 2, batch 7 (1 GPU): pinned prefetch loads + dynamic groups in the pair kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "pair or fresh" --tb=short 2>&1 | tail -4
run() {   # name, bench args
  timeout 200 python bench.py --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'd2h', d['e2e']['d2h_bytes_per_step'], 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | tail -2
}
run b7_default ""
run b7_fe4 "--pair-flush-every 4"
run b7_skew "--composition skewed"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/r02_launches_b7.csv python bench.py --workload count --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/r02_launches_b7.csv')):
    if len(r) > 5 and ('pair_' in r[4] or 'finalize' in r[4] or 'emset' in r[4]): print(r[4][:50], r[-1])
PY
