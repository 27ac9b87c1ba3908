#!/bin/bash
# round 2 (8 GPUs, charged 8x): count bench end to end with the hybrid upload under torchrun
mkdir -p gpurun_out
nproc
run() {
  KPAL_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --workload count --steps 10 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
e=d['e2e']
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'e2e', round(e['value'],1), round(e['ms_per_step'],3), 'h2d', e['h2d_bytes_per_step'], 'host_frac', e.get('host_packed_text_frac'), 'parity', d['parity_ok'])"
  grep "kpal trace" gpurun_out/r02_$1.err | tail -3
  grep -v "^$" gpurun_out/r02_$1.err | grep -v "Warning\|kpal trace\|OMP_NUM\|\*\*\*" | tail -2
}
run c8_hybrid ""
run c8_hybrid_off "--fasta-hybrid 0"
