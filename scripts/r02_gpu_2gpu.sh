#!/bin/bash
# round 2, 2-GPU session (charged 2x): NCCL tests, then the default bench (count + matrix) at N = 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "distributed_nccl" --tb=short 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02_bench_2gpu.json').read().strip().splitlines()[-1])
    print('count N=2: ms/step', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 2),
          round(d['e2e']['ms_per_step'], 3), 'parity', d['parity_ok'], d['parity'])
    m = d['matrix']
    print('matrix N=2: ms/step', round(m['ms_per_step'], 1), 'value', round(m['value']), 'e2e', m['e2e'] and round(m['e2e']['value']),
          m['e2e'] and round(m['e2e']['ms_per_step']), 'parity', m['parity_ok'], m['parity'])
    print(m['config']['parallelism'])
except Exception as exc:
    print('no line:', exc)
PY
grep -v "^W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/r02_bench_2gpu.err | tail -12
