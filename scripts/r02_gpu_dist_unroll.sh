#!/bin/bash
# round 2: distance tile kernel, inner-loop unroll variants at 2048 profiles
mkdir -p gpurun_out
run() {   # name, env, args
  env $2 timeout 600 python bench.py --workload matrix --steps 2 --warmup 1 --no-gram --no-e2e --profiles 2048 > gpurun_out/r02_dist_$1.json 2> gpurun_out/r02_dist_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_dist_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],1), 'value', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_dist_$1.err | tail -2
}
run u2 "A=1"
run u1 "KPAL_B200_LIB=$PWD/kpal_b200/libkpal_b200_u1.so"
run u4 "KPAL_B200_LIB=$PWD/kpal_b200/libkpal_b200_u4.so"
run u2b "A=1"
