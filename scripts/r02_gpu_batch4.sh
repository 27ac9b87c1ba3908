#!/bin/bash
# round 2, batch 4 (1 GPU): sliced reduce on virtual ranks, pair path with the spread in-kernel memset
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "slice or virtual" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "pair or fresh or tiled" 2>&1 | tail -4
run() {   # name, bench args
  timeout 300 python bench.py --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | tail -2
}
run b4_default ""
run b4_memset "--fresh 0"
run b4_skew "--composition skewed"
run b4_cfg1 "--config 1"
run b4_cfg5 "--config 5 --steps 5"
