#!/bin/bash
# 8-GPU count bench with the default flags (charged 8x: one short run)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_count_8gpu.json 2> gpurun_out/bench_count_8gpu.err
tail -1 gpurun_out/bench_count_8gpu.json | cut -c1-1500; grep -v "^W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_count_8gpu.err | tail -4
