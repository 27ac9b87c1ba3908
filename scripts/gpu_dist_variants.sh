#!/bin/bash
# distance tile kernel geometry variants at 2048 profiles (k=10, scaled multiset/prod)
mkdir -p gpurun_out
for v in ${VARIANTS:-"" _r232 _r200}; do
  lib=kpal_b200/libkpal_b200$v.so
  [ -f $lib ] || continue
  KPAL_B200_LIB=$PWD/$lib timeout 600 python bench.py --workload matrix --profiles ${NPROF:-2048} --steps 2 --warmup 1 2>gpurun_out/variant_err$v.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('variant [$v]', 'ms', round(d['ms_per_step'], 1), 'pairs/s', round(d['value']), 'frac', round(d['roofline']['frac'], 4), 'relerr', d['max_rel_err_vs_oracle_276_pairs'], d['parity_ok'])"
done
