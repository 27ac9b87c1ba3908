#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3 4 8 16; do
timeout 600 python bench.py --steps 20 --warmup 3 --fasta-chunks $c 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('cfg2 chunks $c', 'e2e_ms', round(d['e2e']['ms_per_step'], 3), 'e2e', round(d['e2e']['value'],2), 'parity', d['parity_ok'])"
done
