#!/bin/bash
# 4-GPU matrix bench (charged 4x: one short run)
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 4 --workload matrix --steps 2 --warmup 1 > gpurun_out/bench_matrix_4gpu.json 2> gpurun_out/bench_matrix_4gpu.err
tail -1 gpurun_out/bench_matrix_4gpu.json | cut -c1-900; grep -v "^W\|^$\|\*\*\*\|OMP_NUM" gpurun_out/bench_matrix_4gpu.err | tail -4
