#!/bin/bash
# ncu --set full of the pair kernels inside the count bench (one launch each)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_ -s 9 -c 3 \
    -o gpurun_out/r02_ncu_pairs -f python bench.py --workload count --steps 2 --warmup 3 --pair-upt ${UPT:-1} > gpurun_out/ncu_pairs.log 2>&1
tail -3 gpurun_out/ncu_pairs.log
ls -la gpurun_out/r02_ncu_pairs.ncu-rep
