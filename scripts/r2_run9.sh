#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --import-source on -k regex:knn_tc_kernel -c 1 -s 1 -f -o /tmp/knn_tc python scripts/knn_one.py > gpurun_out/ncu_knn_tc.log 2>&1
ncu -i /tmp/knn_tc.ncu-rep --page details 2>/dev/null | grep -E "Duration|Elapsed Cycles|SM Frequency|Issue Slots Busy|Executed Ipc|No Eligible|Eligible Warps|Active Warps|Warp Cycles Per Issued|Stall|L2 Cache Throughput|Memory Throughput|Registers|Achieved Occupancy|Theoretical Occ|Executed Instructions  " | head -40
python scripts/ncu_source_mix.py /tmp/knn_tc.ncu-rep 0 2>&1 | head -30
ncu -i /tmp/knn_tc.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/knn_tc_source.csv.gz
