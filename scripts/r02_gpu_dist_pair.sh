#!/bin/bash
# round 2: distance tile kernel with one reciprocal per two terms: parity tests, then A/B at 2048 profiles and the 4096-profile record
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distance.py -x -q -m gpu --tb=short 2>&1 | tail -4
run() {   # name, env, args
  env $2 timeout 600 python bench.py --workload matrix --steps 2 --warmup 1 --no-gram $3 > gpurun_out/r02_dist_$1.json 2> gpurun_out/r02_dist_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_dist_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity_ok'], d['parity'].get('max_rel_err_leading_block'), d['parity'].get('max_rel_err_random_pairs'))"
  grep -v "^$" gpurun_out/r02_dist_$1.err | tail -2
}
run pair2048 "A=1" "--profiles 2048 --no-e2e"
run rcp1_2048 "KPAL_B200_LIB=$PWD/kpal_b200/libkpal_b200_rcp1.so" "--profiles 2048 --no-e2e"
run pair4096 "A=1" "--no-e2e"
