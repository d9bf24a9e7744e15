#!/bin/bash
# round 2 profiling pass (1 GPU): default bench, ncu launch list of the same command, ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 600 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_partial_h3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-workloads > gpurun_out/ncu_launch.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_partial_h3.csv > gpurun_out/launches_partial_h3_summary.txt
head -40 gpurun_out/launches_partial_h3_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:gemm_tc|flash_attn|knn_select|knn3_kernel|edgeconv_dg_tc|attn_colsum|softcorr_tc|layernorm_operand|select_stats|layernorm_kernel' -s 160 -c 44 \
    -f -o /tmp/prof_h3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-workloads > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_h3.ncu-rep --page raw --csv > gpurun_out/prof_h3_raw.csv 2>> gpurun_out/ncu_full.log
python scripts/summarize_ncu_full.py gpurun_out/prof_h3_raw.csv > gpurun_out/ncu_full_summary.txt
python scripts/ncu_traffic.py gpurun_out/prof_h3_raw.csv > gpurun_out/ncu_traffic.json
head -50 gpurun_out/ncu_full_summary.txt
rm -f gpurun_out/prof_h3_raw.csv.gz; gzip -9 -k gpurun_out/prof_h3_raw.csv 2>/dev/null; rm -f gpurun_out/prof_h3_raw.csv
timeout 600 python scripts/kernel_sweep.py > gpurun_out/kernel_sweep_cfg5.txt 2>&1; tail -30 gpurun_out/kernel_sweep_cfg5.txt
