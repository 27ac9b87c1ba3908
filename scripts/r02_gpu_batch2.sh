#!/bin/bash
# round 2, batch 2: new tests + A/B of the pair path variants + the default bench (both legs) + smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m pytest tests/test_gpu_count.py tests/test_gpu_distance.py tests/test_multigpu.py -x -q -m gpu \
   -k "pair or fresh or tiled or 600 or concurrent or 20k or virtual or narrow" 2>&1 | tail -8
run() {   # name, bench args
  timeout 300 python bench.py --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | tail -2
}
run b2_default ""
run b2_fe4 "--pair-flush-every 4"
run b2_fe2 "--pair-flush-every 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r02_launches_b2.csv python bench.py --workload count --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/r02_launches_b2.csv')):
    if len(r) > 5 and ('pair_' in r[4] or 'finalize' in r[4] or 'emset' in r[4]): print(r[4][:50], r[-1])
PY
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -c 3000 gpurun_out/r02_bench_default.json; grep -v "^$" gpurun_out/r02_bench_default.err | tail -3
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
