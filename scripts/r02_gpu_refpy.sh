#!/bin/bash
# round 2: the reference arm with the unmodified reference's Python beside the C port (baseline/_ref travels with the snapshot)
mkdir -p gpurun_out
ls baseline/_ref | head -3
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_python.json 2> gpurun_out/r02_bench_reference_python.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_reference_python.json').read().strip().splitlines()[-1])
print(d['value'], d['cpu_baseline']['cores'], d['reference_python']); print(d.get('matrix',{}).get('value'))"
tail -2 gpurun_out/r02_bench_reference_python.err
