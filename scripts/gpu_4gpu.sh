#!/bin/bash
# 4-GPU session (charged 4x: keep it short): the count bench with the table sum fused into
# the count over NVLink peer memory (the default at N > 1), bit-checked against the NCCL sum
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_count_4gpu_fused.json 2> gpurun_out/bench_count_4gpu_fused.err
tail -1 gpurun_out/bench_count_4gpu_fused.json; grep -v "^W\|^$" gpurun_out/bench_count_4gpu_fused.err | tail -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 4 --steps 10 --warmup 3 --reduce nccl > gpurun_out/bench_count_4gpu_nccl.json 2> gpurun_out/bench_count_4gpu_nccl.err
tail -1 gpurun_out/bench_count_4gpu_nccl.json; grep -v "^W\|^$" gpurun_out/bench_count_4gpu_nccl.err | tail -5
