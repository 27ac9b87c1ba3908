#!/bin/bash
# per-record counting (BASELINE configs[2]): tests, then the bench (device-resident + host entry point)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "by_record" 2>&1 | tail -5
free -g | head -2
timeout 900 python scripts/bench_by_record.py 2>&1 | tail -4 | tee gpurun_out/bench_by_record_slab.jsonl
