#!/bin/bash
# round 2: by-record output path (sparse deflate, fast row stats, reused buffers + overlap): tests, then the file-to-file run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py tests/test_gpu_cli.py -x -q -m gpu -k "by_record or command_line" --tb=short 2>&1 | tail -4
timeout 900 python scripts/bench_by_record_cli.py > gpurun_out/r02_by_record_cli2.jsonl 2> gpurun_out/r02_by_record_cli2.err
cat gpurun_out/r02_by_record_cli2.jsonl; grep -v "^$" gpurun_out/r02_by_record_cli2.err | tail -3

KPAL_BY_RECORD_N=100000 timeout 900 python - <<'PY' 2>&1 | tail -25
import cProfile, pstats, os, sys, runpy
sys.argv = ['scripts/bench_by_record_cli.py']
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path('scripts/bench_by_record_cli.py', run_name='__main__')
finally:
    pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
PY
