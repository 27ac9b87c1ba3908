#!/bin/bash
# quick A/B of the count paths (value line only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "radix or tiled" 2>&1 | tail -5
for args in "--count-path 1" "--count-path 2 --radix-shape 1" "--count-path 2 --radix-shape 2"; do
  timeout 300 python bench.py --steps 10 --warmup 3 $args 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$args', 'step_ms', round(d['ms_per_step'], 4), 'kern_ms', round(d['roofline']['kernel_ms'], 4), 'e2e_ms', round(d['e2e']['ms_per_step'], 3), 'parity', d['parity_ok'])"
done
for sh in 1 2; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/l.csv python bench.py --steps 2 --warmup 3 --count-path 2 --radix-shape $sh > gpurun_out/ncu_bench.log 2>&1
echo "shape $sh: partition ns:" $(grep -E "radix_partition" gpurun_out/l.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -3 | tr '\n' ' ') "hist ns:" $(grep -E "radix_histogram" gpurun_out/l.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -3 | tr '\n' ' ')
done
