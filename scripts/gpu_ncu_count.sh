#!/bin/bash
# ncu --set full of the radix count kernels (one launch each, after warm-up)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'radix_partition_kernel|radix_histogram_kernel' -s 6 -c 2 \
    -o gpurun_out/ncu_count_radix -f python bench.py --steps 2 --warmup 3 --count-path 2 ${EXTRA} > gpurun_out/ncu_count_radix.log 2>&1
tail -3 gpurun_out/ncu_count_radix.log
ncu -i gpurun_out/ncu_count_radix.ncu-rep --page details > gpurun_out/ncu_count_radix_details.txt
ls -la gpurun_out/
