#!/bin/bash
# ncu evidence for the count path: launch list of the bench + --set full of the two radix kernels
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench_count.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'radix_partition_kernel|radix_histogram_kernel|finalize_balance_tiled' -s 9 -c 3 \
    -o gpurun_out/ncu_count_radix -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_count_radix.log 2>&1
tail -3 gpurun_out/ncu_count_radix.log
ncu -i gpurun_out/ncu_count_radix.ncu-rep --page details > gpurun_out/ncu_count_radix_details.txt
ncu -i gpurun_out/ncu_count_radix.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_red.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum > gpurun_out/ncu_count_radix_raw.csv
