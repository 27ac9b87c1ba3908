#!/bin/bash
# end-to-end count: the call's host + device timeline (KPAL_TRACE) per upload chunking
mkdir -p gpurun_out
run() {   # name, bench args
  KPAL_TRACE=1 timeout 300 python bench.py --steps 10 $2 > gpurun_out/bench_count_$1.json 2> gpurun_out/bench_count_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_count_$1.json').read().strip().splitlines()[-1])
print('$1', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), d['e2e']['d2h_bytes_per_step'], 'parity', d['parity_ok'])"
  grep "kpal trace" gpurun_out/bench_count_$1.err | tail -2
  grep -v "kpal trace" gpurun_out/bench_count_$1.err | tail -2
}
run chunks16 "--fasta-chunks 16"
run chunks12 "--fasta-chunks 12"
run chunks8 "--fasta-chunks 8"
run chunks6 "--fasta-chunks 6"
run chunks4 "--fasta-chunks 4"
run chunks2 "--fasta-chunks 2"
