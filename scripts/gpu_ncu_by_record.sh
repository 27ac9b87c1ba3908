#!/bin/bash
# ncu --set full of by_record_kernel (shared-memory slabs) in the by-record bench (20 k records per launch)
mkdir -p gpurun_out
KPAL_BY_RECORD_E2E=0 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'by_record_kernel' -s 5 -c 1 -o gpurun_out/ncu_by_record -f python scripts/bench_by_record.py > gpurun_out/ncu_by_record.log 2>&1
tail -3 gpurun_out/ncu_by_record.log
ncu -i gpurun_out/ncu_by_record.ncu-rep --page details > gpurun_out/ncu_by_record_details.txt
ncu -i gpurun_out/ncu_by_record.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed_op_shared_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active > gpurun_out/ncu_by_record_raw.csv
cat gpurun_out/ncu_by_record_raw.csv | tail -3
grep -E "Duration|DRAM Throughput|Memory Throughput|Achieved Occupancy|Registers Per|Dynamic Shared" gpurun_out/ncu_by_record_details.txt | head -12
