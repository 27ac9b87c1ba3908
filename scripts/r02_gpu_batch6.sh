#!/bin/bash
# round 2, batch 6 (1 GPU): tests of the latest changes, full config-3 CLI run, ncu captures for profiles/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_gpu_count.py tests/test_gpu_cli.py -x -q -m gpu \
   -k "slice or narrow_profile or cli or command or by_record" --tb=short 2>&1 | tail -8
timeout 900 python scripts/bench_by_record_cli.py 2>&1 | tail -3 | tee gpurun_out/r02_by_record_cli.jsonl
timeout 900 python scripts/bench_by_record.py 2>&1 | tail -4 | tee gpurun_out/r02_by_record.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pair_|finalize_balance" -s 12 -c 3 \
    -o gpurun_out/r02_ncu_count -f python bench.py --workload count --steps 2 --warmup 3 > gpurun_out/ncu_count.log 2>&1
tail -2 gpurun_out/ncu_count.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_u8 -s 1 -c 1 \
    -o gpurun_out/r02_ncu_gram -f python bench.py --workload matrix --profiles 2048 --steps 1 --no-e2e > gpurun_out/ncu_gram.log 2>&1
tail -2 gpurun_out/ncu_gram.log | cut -c1-300
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:distance_tile -c 1 --csv --log-file gpurun_out/r02_ncu_distance_tile_4096_dram.csv \
    python bench.py --workload matrix --steps 1 --no-e2e --no-gram > gpurun_out/ncu_dist.log 2>&1
tail -3 gpurun_out/r02_ncu_distance_tile_4096_dram.csv | cut -c1-400
ls -la gpurun_out/*.ncu-rep
