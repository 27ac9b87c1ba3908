#!/bin/bash
# after the last library change: multi-GPU driver + CLI tests, smoke, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_gpu_cli.py tests/test_gpu_split.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
