#!/bin/bash
# round 2, 8-GPU session b (charged 8x): count at N = 8 with the tile-major push, with / without the NUMA binding
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open('gpurun_out/%s.json' % name).read().strip().splitlines()[-1])
    print(name, 'count: ms/step', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'kernel_ms', round(d['roofline']['count_kernels_ms'], 4),
          'tail', {k: round(v, 4) for k, v in (d.get('reduce_tail') or {}).items() if k != 'note'},
          'e2e', round(d['e2e']['value'], 2), round(d['e2e']['ms_per_step'], 3), 'parity', d['parity_ok'])
except Exception as exc:
    print(name, 'no line:', exc)
PY
grep -v "^W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/$1.err | tail -4
}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 8 --workload count --steps 10 --warmup 3 > gpurun_out/r02_count_8gpu_bind.json 2> gpurun_out/r02_count_8gpu_bind.err
show r02_count_8gpu_bind
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus 8 --workload count --steps 10 --warmup 3 --numa-bind 0 > gpurun_out/r02_count_8gpu_nobind.json 2> gpurun_out/r02_count_8gpu_nobind.err
show r02_count_8gpu_nobind
python - <<'PY'
import os
try:
    import pynvml
    pynvml.nvmlInit()
    for i in range(pynvml.nvmlDeviceGetCount()):
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        print('gpu', i, 'ideal cpus', cpus[0], '..', cpus[-1], len(cpus))
    print('process affinity', len(os.sched_getaffinity(0)), 'cpu_count', os.cpu_count())
except Exception as exc:
    print('nvml:', exc)
PY
