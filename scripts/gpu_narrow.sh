#!/bin/bash
# narrow (uint16) D2H of the profile: count tests, then the bench with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -5
for n in 1 0; do
  timeout 300 python bench.py --narrow-d2h $n > gpurun_out/bench_count_narrow$n.json 2> gpurun_out/bench_count_narrow$n.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_count_narrow$n.json').read().strip().splitlines()[-1])
print('narrow=$n', 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['parity_ok'])"
  tail -2 gpurun_out/bench_count_narrow$n.err
done
timeout 300 python bench.py --config 5 --steps 5 > gpurun_out/bench_count_cfg5_narrow1.json 2> gpurun_out/bench_count_cfg5.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_count_cfg5_narrow1.json').read().strip().splitlines()[-1])
print('cfg5', 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d['parity_ok'])"
