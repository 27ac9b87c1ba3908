#!/bin/bash
# 2-GPU session: multi-GPU tests, then the count bench with the table sum fused into the
# count (default), as separate peer-memory kernels, and with NCCL; then the matrix bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -6
timeout 600 python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -8
for mode in fused peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --reduce $mode > gpurun_out/bench_count_2gpu_$mode.json 2> gpurun_out/bench_count_2gpu_$mode.err
  tail -1 gpurun_out/bench_count_2gpu_$mode.json; grep -v "^W\|^$" gpurun_out/bench_count_2gpu_$mode.err | tail -5
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 5 --warmup 3 --config 5 --mbp-per-gpu 375 > gpurun_out/bench_count_cfg5_2gpu.json 2> gpurun_out/bench_count_cfg5_2gpu.err
tail -1 gpurun_out/bench_count_cfg5_2gpu.json; grep -v "^W\|^$" gpurun_out/bench_count_cfg5_2gpu.err | tail -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --workload matrix --steps 2 --warmup 1 > gpurun_out/bench_matrix_2gpu.json 2> gpurun_out/bench_matrix_2gpu.err
tail -1 gpurun_out/bench_matrix_2gpu.json; grep -v "^W\|^$" gpurun_out/bench_matrix_2gpu.err | tail -5
