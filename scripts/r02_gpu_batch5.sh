#!/bin/bash
# round 2, batch 5 (1 GPU): count tests after the clean-up, narrow side lists, skewed composition
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "pair or fresh or narrow" 2>&1 | tail -4
run() {   # name, bench args
  timeout 200 python bench.py --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'd2h', d['e2e']['d2h_bytes_per_step'], 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | tail -2
}
run b5_default ""
run b5_skew "--composition skewed"
