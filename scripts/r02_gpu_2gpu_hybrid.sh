#!/bin/bash
# round 2 (2 GPUs): multi-GPU tests + count bench end to end with the hybrid upload under torchrun
mkdir -p gpurun_out
nproc
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu --tb=short 2>&1 | tail -4
run() {
  KPAL_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
e=d['e2e']
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'e2e', round(e['value'],1), round(e['ms_per_step'],3), 'h2d', e['h2d_bytes_per_step'], 'host_frac', e.get('host_packed_text_frac'), 'parity', d['parity_ok'])"
  grep "kpal trace" gpurun_out/r02_$1.err | tail -2
  grep -v "^$" gpurun_out/r02_$1.err | grep -v "Warning\|kpal trace\|OMP_NUM" | tail -2
}
run c2_hybrid ""
run c2_hybrid_off "--fasta-hybrid 0"
