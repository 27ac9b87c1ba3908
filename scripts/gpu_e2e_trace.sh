#!/bin/bash
# end-to-end count: the narrow-copy tests, then the bench per copy width with the call trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -5
for n in 1 2; do
  KPAL_TRACE=1 timeout 300 python bench.py --narrow-d2h $n > gpurun_out/bench_count_narrow$n.json 2> gpurun_out/bench_count_narrow$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_count_narrow$n.json').read().strip().splitlines()[-1])
print('narrow=$n', 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['d2h_bytes_per_step'], 'parity', d['parity_ok'])"
  grep "kpal trace" gpurun_out/bench_count_narrow$n.err | tail -4
  grep -v "kpal trace" gpurun_out/bench_count_narrow$n.err | tail -2
done
nproc; lscpu | grep -E "Model name|Socket|Thread|Core" 
