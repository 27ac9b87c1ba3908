#!/bin/bash
# round 2, batch 11 (1 GPU): ncu of the slice push / collect kernels (virtual world on one device)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slice_ -c 40 -f -o gpurun_out/r02_ncu_slice \
  python scripts/r02_slice_tail.py 12 1 > gpurun_out/r02_ncu_slice.log 2>&1
tail -3 gpurun_out/r02_ncu_slice.log
