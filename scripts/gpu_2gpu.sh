#!/bin/bash
# 2-GPU session: multi-GPU tests, then the count bench with the peer-memory reduce and with NCCL
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -8
for mode in peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --reduce $mode > gpurun_out/bench_count_2gpu_$mode.json 2> gpurun_out/bench_count_2gpu_$mode.err
  tail -1 gpurun_out/bench_count_2gpu_$mode.json; grep -v "^W\|^$" gpurun_out/bench_count_2gpu_$mode.err | tail -5
done
