#!/bin/bash
# round 2: hybrid FASTA upload (host threads pack the tail of the text while the head uploads raw)
mkdir -p gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "hybrid" --tb=short 2>&1 | tail -4
run() {   # name, bench args
  KPAL_TRACE=1 timeout 300 python bench.py --workload count --steps 10 $2 > gpurun_out/r02_hyb_$1.json 2> gpurun_out/r02_hyb_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02_hyb_$1.json').read().strip().splitlines()[-1])
e=d['e2e']
print('$1', 'value', round(d['value'],1), 'e2e', round(e['value'],2), round(e['ms_per_step'],3), 'h2d', e['h2d_bytes_per_step'], 'host_frac', e.get('host_packed_text_frac'), 'parity', d['parity_ok'])"
  grep "kpal trace" gpurun_out/r02_hyb_$1.err | tail -${3:-2}
  grep -v "kpal trace" gpurun_out/r02_hyb_$1.err | tail -2
}
run early20 "--fasta-hybrid 1 --steps 20" 3
run early10 "--fasta-hybrid 1" 2
run cfg5early "--config 5" 2
run skewed_early "--composition skewed" 2
