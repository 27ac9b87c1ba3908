#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 5 7; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/l.csv python bench.py --steps 2 --warmup 3 --count-path 2 --radix-debug $dbg > gpurun_out/ncu_bench.log 2>&1
echo "debug $dbg: partition us:" $(grep -E "radix_partition" gpurun_out/l.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -3 | tr '\n' ' ')
done
