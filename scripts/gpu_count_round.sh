#!/bin/bash
# One GPU session for the counting path: parity of the radix path, A/B bench, launch list.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "radix or tiled" 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --count-path 1 > gpurun_out/bench_count_red.json 2> gpurun_out/bench_count_red.err
cat gpurun_out/bench_count_red.json
timeout 300 python bench.py --steps 10 --warmup 3 --count-path 2 > gpurun_out/bench_count_radix.json 2> gpurun_out/bench_count_radix.err
cat gpurun_out/bench_count_radix.json; tail -3 gpurun_out/bench_count_radix.err
timeout 300 python bench.py --steps 10 --warmup 3 --count-path 2 --radix-payload-bits 14 > gpurun_out/bench_count_radix_p14.json 2>&1
cat gpurun_out/bench_count_radix_p14.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_count_radix.csv python bench.py --steps 2 --warmup 3 --count-path 2 > gpurun_out/ncu_bench.log 2>&1
grep -E "radix|finalize|fasta" gpurun_out/launches_count_radix.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -30
