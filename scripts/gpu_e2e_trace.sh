#!/bin/bash
# end-to-end count: the two-part upload/count overlap -- tests, then the bench with the call trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_count.py -x -q -m gpu -k "two_overlapped or chunked_upload or narrow or host_api" 2>&1 | tail -5
run() {   # name, bench args
  KPAL_TRACE=1 timeout 300 python bench.py --steps 10 $2 > gpurun_out/bench_count_$1.json 2> gpurun_out/bench_count_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/bench_count_$1.json').read().strip().splitlines()[-1])
print('$1', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), d['e2e']['d2h_bytes_per_step'], 'parity', d['parity_ok'])"
  grep "kpal trace" gpurun_out/bench_count_$1.err | tail -2
  grep -v "kpal trace" gpurun_out/bench_count_$1.err | tail -2
}
run split1 "--fasta-split 1"
run split0 "--fasta-split 0"


