#!/bin/bash
# round 2, final 1-GPU session: whole GPU test suite, smoke, default bench + reference arm, the other count
# configurations, launch lists and ncu --set full captures of the count step and the slice kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('default count: ms/step', round(d['ms_per_step'], 4), 'value', round(d['value'], 1), 'frac', round(d['roofline']['frac'], 3),
      'e2e', round(d['e2e']['value'], 2), round(d['e2e']['ms_per_step'], 3), 'launches', d['gpu_launches'], 'parity', d['parity_ok'], 'clocks', d['clocks'])
m = d['matrix']
print('default matrix: ms/step', round(m['ms_per_step'], 1), 'value', round(m['value']), 'frac', round(m['roofline']['frac'], 3),
      'e2e', m['e2e'] and round(m['e2e']['value']), 'parity', m['parity_ok'])
g = m.get('euclidean_gram')
print('gram', g and {k: g[k] for k in g if k in ('value', 'ms_per_step', 'parity_ok', 'speedup_vs_tile_kernel')})
print('cpu', d['cpu_baseline'])
PY
tail -2 gpurun_out/r02_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-600 gpurun_out/r02_bench_reference.json
run() {   # name, bench args
  timeout 300 python bench.py --workload count --steps 20 $2 > gpurun_out/r02_$1.json 2> gpurun_out/r02_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['count_kernels_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'd2h', d['e2e']['d2h_bytes_per_step'], 'parity', d['parity_ok'])"
  grep -v "^$" gpurun_out/r02_$1.err | tail -2
}
run bench_count_skewed "--composition skewed"
run bench_count_cfg1 "--config 1"
run bench_count_cfg5_shard "--config 5 --steps 10"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r02_launches_bench_count.csv python bench.py --workload count --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pair_|finalize_balance" -s 6 -c 3 \
    -o gpurun_out/r02_ncu_count_final -f python bench.py --workload count --steps 1 --warmup 3 > gpurun_out/ncu_count_final.log 2>&1
tail -2 gpurun_out/ncu_count_final.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slice_ -c 6 -f -o gpurun_out/r02_ncu_slice_final \
  python scripts/r02_slice_tail.py 12 1 > gpurun_out/r02_ncu_slice_final.log 2>&1
tail -2 gpurun_out/r02_ncu_slice_final.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/r02_launches_bench_matrix_n512.csv python bench.py --workload matrix --profiles 512 --steps 1 --warmup 1 > gpurun_out/ncu_bench_m.log 2>&1
tail -2 gpurun_out/ncu_bench_m.log
