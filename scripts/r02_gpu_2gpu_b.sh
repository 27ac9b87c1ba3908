#!/bin/bash
# round 2 (2 GPUs): sliced reduce after the push/collect rewrite: tests, count bench with the tail split, matrix
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu --tb=short 2>&1 | tail -4
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'tail', d.get('reduce_tail'), 'e2e', round(d['e2e']['value'],1), 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | grep -v Warning | tail -2
}
timeout 200 python scripts/r02_slice_tail.py 12 2>&1 | tail -4
run c2_slices ""
run c2_slices_again ""
