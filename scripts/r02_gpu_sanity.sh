#!/bin/bash
# round 2: last sanity pass after moving the hybrid upload's adaptation state into the per-device workspace
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "hybrid or full_size or exotic or by_record_many" --tb=short 2>&1 | tail -3
KPAL_TRACE=1 timeout 300 python bench.py --workload count > gpurun_out/r02_sanity_count.json 2> gpurun_out/r02_sanity_count.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_sanity_count.json').read().strip().splitlines()[-1])
e=d['e2e']
print('count', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'e2e', round(e['value'],2), round(e['ms_per_step'],3), 'h2d', e['h2d_bytes_per_step'], 'host_frac', e.get('host_packed_text_frac'), 'parity', d['parity_ok'])"
grep "kpal trace" gpurun_out/r02_sanity_count.err | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
