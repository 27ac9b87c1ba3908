#!/bin/bash
# round 2, batch 9 (1 GPU): the multi-GPU tail kernels alone (virtual world on one device) + ncu of them
mkdir -p gpurun_out
timeout 300 python scripts/r02_slice_tail.py 12 2>&1 | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slice_ -c 8 -f -o gpurun_out/r02_ncu_slice \
  python scripts/r02_slice_tail.py 12 > gpurun_out/r02_ncu_slice.log 2>&1
tail -3 gpurun_out/r02_ncu_slice.log
