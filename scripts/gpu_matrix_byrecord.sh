#!/bin/bash
# matrix bench (device-resident + end-to-end legs) and the per-record bench (BASELINE configs[2])
mkdir -p gpurun_out
free -g | head -2
timeout 900 python bench.py --workload matrix --steps 2 --warmup 1 > gpurun_out/bench_matrix.json 2> gpurun_out/bench_matrix.err
cat gpurun_out/bench_matrix.json; tail -5 gpurun_out/bench_matrix.err
timeout 600 python scripts/bench_by_record.py > gpurun_out/bench_by_record.jsonl 2> gpurun_out/bench_by_record.err
cat gpurun_out/bench_by_record.jsonl; tail -3 gpurun_out/bench_by_record.err
