#!/bin/bash
# round 2: by-record host entry point with workspace buffers + finer D2H / widen pipelining
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py tests/test_gpu_cli.py -x -q -m gpu -k "by_record or command_line" --tb=short 2>&1 | tail -3
timeout 900 python scripts/bench_by_record_cli.py > gpurun_out/r02_by_record_cli3.jsonl 2> gpurun_out/r02_by_record_cli3.err
cat gpurun_out/r02_by_record_cli3.jsonl; grep -v "^$" gpurun_out/r02_by_record_cli3.err | tail -3
timeout 900 python scripts/bench_by_record.py 2>&1 | tail -4 | tee gpurun_out/r02_by_record3.jsonl
