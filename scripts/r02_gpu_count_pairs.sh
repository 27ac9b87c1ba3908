#!/bin/bash
# round 2: the two-windows-per-payload count path -- parity tests, A/B of its variants
# against the one-window radix path, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "pair" 2>&1 | tail -5
run() {   # name, bench args
  timeout 300 python bench.py --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | tail -2
}
run pairs_default ""
run pairs_fe1 "--pair-flush-every 1"
run pairs_fe2 "--pair-flush-every 2"
run pairs_fe6 "--pair-flush-every 6"
run pairs_upt2 "--pair-upt 2"
run pairs_unfused "--pair-fused 0"
run onewin "--count-path 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/r02_launches_pairs.csv python bench.py --workload count --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
grep -E "pair_|finalize|memset|Memset" gpurun_out/r02_launches_pairs.csv | awk -F'","' '{print $5, $NF}' | tr -d '"' | cut -c1-60,100- | tail -8
