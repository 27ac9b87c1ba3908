#!/bin/bash
# round 2, batch 10 (1 GPU): push kernel with one fence per CTA + dynamic tiles: virtual-world parity and timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu -k "virtual" --tb=short 2>&1 | tail -4
timeout 300 python scripts/r02_slice_tail.py 12 2>&1 | tail -8
timeout 300 python scripts/r02_slice_tail.py 10 2>&1 | tail -8
