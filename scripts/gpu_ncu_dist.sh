#!/bin/bash
# ncu --set full of the distance tile kernel (one launch, 512 profiles)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'distance_tile_kernel' -s 1 -c 1 \
    -o gpurun_out/ncu_dist -f python bench.py --workload matrix --profiles ${NPROF:-512} --steps 1 --warmup 1 > gpurun_out/ncu_dist.log 2>&1
tail -2 gpurun_out/ncu_dist.log
ncu -i gpurun_out/ncu_dist.ncu-rep --page details > gpurun_out/ncu_dist_details.txt
