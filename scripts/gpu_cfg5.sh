#!/bin/bash
# config 5 shard (k=13, 375 Mbp per GPU): buckets per pass-1 launch 1024 (two sweeps) vs 2048 (one)
mkdir -p gpurun_out
for mb in 1024 2048; do
  timeout 400 python bench.py --config 5 --steps 5 --radix-max-buckets $mb > gpurun_out/bench_cfg5_mb$mb.json 2> gpurun_out/bench_cfg5_mb$mb.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg5_mb$mb.json').read().strip().splitlines()[-1])
print('max_buckets $mb', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'parity', d['parity_ok'])"
  tail -2 gpurun_out/bench_cfg5_mb$mb.err
done
