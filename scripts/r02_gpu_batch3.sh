#!/bin/bash
# round 2, batch 3: Gram (tcgen05) tests first, pair path A/B (in-kernel memset on/off), matrix bench with the Gram leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distance.py -x -q -m gpu -k "gram or concurrent" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "pair or fresh" 2>&1 | tail -4
run() {   # name, bench args
  timeout 300 python bench.py --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | tail -2
}
run b3_default ""
run b3_memset "--fresh 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/r02_launches_b3.csv python bench.py --workload count --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/r02_launches_b3.csv')):
    if len(r) > 5 and ('pair_' in r[4] or 'finalize' in r[4] or 'emset' in r[4]): print(r[4][:50], r[-1])
PY
timeout 900 python bench.py --workload matrix --profiles 1024 --steps 2 --no-e2e > gpurun_out/r02_matrix1024.json 2> gpurun_out/r02_matrix1024.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_matrix1024.json').read().strip().splitlines()[-1])
print('matrix1024', d['ms_per_step'], d['parity'], d['euclidean_gram'])"
grep -v "^$" gpurun_out/r02_matrix1024.err | tail -5
