#!/bin/bash
# round 2, batch 13 (1 GPU): push kernel v3: parity, timing, ncu full with source
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "virtual" --tb=short 2>&1 | tail -4
timeout 300 python scripts/r02_slice_tail.py 12 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slice_ -c 8 -f -o gpurun_out/r02_ncu_slice \
  python scripts/r02_slice_tail.py 12 1 > gpurun_out/r02_ncu_slice.log 2>&1
tail -2 gpurun_out/r02_ncu_slice.log
