#!/bin/bash
# count bench A/B: payload bits of the radix path (15 = 512 buckets, one histogram CTA per SM;
# 14 = 1024 buckets, three per SM) and the upload chunking of the end-to-end leg
mkdir -p gpurun_out
run() {   # name, bench args
  KPAL_TRACE=1 timeout 300 python bench.py --steps 20 $2 > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$1.json').read().strip().splitlines()[-1])
print('$1', 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'parity', d['parity_ok'])"
  grep "kpal trace" gpurun_out/ab_$1.err | tail -1
  grep -v "kpal trace" gpurun_out/ab_$1.err | tail -2
}
run p15_c32 ""
run p14_c32 "--radix-payload-bits 14"
run p15_c16 "--fasta-chunks 16"
run p15_c24 "--fasta-chunks 24"
