#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "radix or tiled" 2>&1 | tail -5
for args in "--count-path 1" "--count-path 2"; do
  timeout 600 python bench.py --config 5 --steps 5 --warmup 3 $args 2>gpurun_out/cfg5.err | tee gpurun_out/bench_cfg5_${args##* }.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$args', d['metric'], round(d['value'],1), 'step_ms', round(d['ms_per_step'], 4), 'kern_ms', round(d['roofline']['kernel_ms'], 4), 'e2e_ms', round(d['e2e']['ms_per_step'], 3), 'e2e', round(d['e2e']['value'],2), 'parity', d['parity_ok'], 'cpu', d['cpu_baseline']['value'])"
  tail -3 gpurun_out/cfg5.err
done
