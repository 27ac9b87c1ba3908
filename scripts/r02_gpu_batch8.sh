#!/bin/bash
# round 2, batch 8 (1 GPU): pass-1 atomics in flight per batch: 8 / 12 / 14 of the 16 pairs of a unit
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "pair or fresh" --tb=short 2>&1 | tail -4
run() {   # name, lib, bench args
  KPAL_B200_LIB=$2 timeout 200 python bench.py --workload count --steps 20 $3 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | tail -2
}
run b8_nb14 "" ""
run b8_nb12 "$PWD/kpal_b200/libkpal_b200_pb12.so" ""
run b8_nb8 "$PWD/kpal_b200/libkpal_b200_pb8.so" ""
run b8_nb14_fe4 "" "--pair-flush-every 4"
