#!/bin/bash
# per-record counting (BASELINE configs[2]): tests, then the bench with the slab kernel and the RED rows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "by_record" 2>&1 | tail -5
timeout 600 python scripts/bench_by_record.py 2>&1 | tail -2 | tee gpurun_out/bench_by_record_slab.jsonl
KPAL_BY_RECORD_PATH=1 timeout 600 python scripts/bench_by_record.py 2>&1 | tail -2 | tee gpurun_out/bench_by_record_red.jsonl
