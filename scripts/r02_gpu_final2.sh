#!/bin/bash
# round 2, final 1-GPU check after the hybrid upload / by-record work: whole GPU test suite, smoke, default bench + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_default_final.json').read().strip().splitlines()[-1])
print('default count: ms/step', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'traffic', d['roofline']['traffic'],
      'e2e', round(d['e2e']['value'], 2), round(d['e2e']['ms_per_step'], 3), d['e2e']['h2d_bytes_per_step'], d['e2e'].get('host_packed_text_frac'), 'launches', d['gpu_launches'], 'parity', d['parity_ok'], 'clocks', d['clocks'])
m = d['matrix']
print('default matrix: ms/step', round(m['ms_per_step'], 1), 'value', round(m['value']), 'frac', round(m['roofline']['frac'], 3), 'traffic', m['roofline']['traffic'],
      'e2e', m['e2e'] and round(m['e2e']['value']), 'parity', m['parity_ok'])
g = m.get('euclidean_gram')
print('gram', g and {k: g[k] for k in g if k in ('value', 'ms_per_step', 'parity_ok')})
print('cpu', d['cpu_baseline'])
PY
tail -2 gpurun_out/r02_bench_default_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err; cut -c1-300 gpurun_out/r02_bench_reference_final.json
