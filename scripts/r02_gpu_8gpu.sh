#!/bin/bash
# round 2, 8-GPU session (charged 8x): default bench (count + matrix) at N = 8, config 5 (human-genome sized, k = 13)
# at N = 8, default count at N = 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
show() {
python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/%s.json' % name).read().strip().splitlines()[-1])
    print(name, 'count: ms/step', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'kernel_ms', round(d['roofline']['count_kernels_ms'], 4),
          'tail', {k: round(v, 4) for k, v in (d.get('reduce_tail') or {}).items() if k != 'note'},
          'e2e', round(d['e2e']['value'], 2), round(d['e2e']['ms_per_step'], 3), 'parity', d['parity_ok'], d['parity'])
    m = d.get('matrix')
    if m:
        print(name, 'matrix: ms/step', round(m['ms_per_step'], 1), 'value', round(m['value']), 'e2e', m['e2e'] and round(m['e2e']['value']),
              m['e2e'] and round(m['e2e']['ms_per_step']), 'parity', m['parity_ok'], m['parity'])
        g = m.get('euclidean_gram')
        if g: print(name, 'gram:', {k: g[k] for k in g if k in ('value', 'ms_per_step', 'parity_ok')})
except Exception as exc:
    print(name, 'no line:', exc)
PY
grep -v "^W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/$1.err | tail -6
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
show r02_bench_8gpu
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 8 --workload count --config 5 --steps 5 --warmup 3 > gpurun_out/r02_cfg5_8gpu.json 2> gpurun_out/r02_cfg5_8gpu.err
show r02_cfg5_8gpu
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus 4 --workload count --steps 10 --warmup 3 > gpurun_out/r02_count_4gpu.json 2> gpurun_out/r02_count_4gpu.err
show r02_count_4gpu
