#!/bin/bash
# new rows: matrix session / formatter / split / showbalance / positive, then the matrix bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_distance.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --workload matrix --steps 2 --warmup 1 > gpurun_out/bench_matrix.json 2> gpurun_out/bench_matrix.err
cat gpurun_out/bench_matrix.json; tail -5 gpurun_out/bench_matrix.err
