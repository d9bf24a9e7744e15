#!/bin/bash
# Artifact pass for profiles/: default bench line (with the CPU arm), launch list, ncu full-set summaries, kernel sweep.
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
cat gpurun_out/bench_reference_arm.json
timeout 600 python bench.py --steps 20 --warmup 3 --workload whole --precision fp16 --no-cpu-baseline > gpurun_out/bench_whole_fp16.json 2> gpurun_out/bench_whole_fp16.err
timeout 300 python scripts/step_profile.py h3 partial > gpurun_out/step_profile_partial_h3.txt 2>&1
timeout 300 python scripts/knn_bench.py > gpurun_out/knn_bench.txt 2>&1
timeout 600 python scripts/kernel_sweep.py > gpurun_out/kernel_sweep_cfg5.txt 2>&1
timeout -s KILL 120 python scripts/flash_diag.py > gpurun_out/flash_diag.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_h3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-workloads > gpurun_out/ncu_launch.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_h3.csv > gpurun_out/launches_h3_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|flash_attn|knn_select|edgeconv_dg_tc|attn_colsum|softcorr_tc' -s 60 -c 28 \
    -f -o /tmp/prof_h3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_h3.ncu-rep --page raw --csv > gpurun_out/prof_h3_raw.csv 2>> gpurun_out/ncu_full.log
python scripts/summarize_ncu_full.py gpurun_out/prof_h3_raw.csv > gpurun_out/ncu_full_summary.txt
python scripts/ncu_traffic.py gpurun_out/prof_h3_raw.csv > gpurun_out/ncu_traffic.json
head -12 gpurun_out/launches_h3_summary.txt
