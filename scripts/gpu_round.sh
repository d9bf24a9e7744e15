#!/bin/bash
# One GPU-box pass: parity tests, bench (default = configs[1] partial; + whole / lpd-train / fp16), per-call profile,
# and (unless "nonc") the ncu launch list + full-set capture of the top kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload whole --no-cpu-baseline > gpurun_out/bench_whole_h3.json 2> gpurun_out/bench_whole_h3.err
cat gpurun_out/bench_whole_h3.json
timeout 600 python bench.py --steps 20 --warmup 3 --workload whole --precision fp16 --no-cpu-baseline > gpurun_out/bench_whole_fp16.json 2> gpurun_out/bench_whole_fp16.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload lpd-train > gpurun_out/bench_lpd_train.json 2> gpurun_out/bench_lpd_train.err
cat gpurun_out/bench_lpd_train.json; tail -3 gpurun_out/bench_lpd_train.err
timeout 300 python scripts/step_profile.py h3 > gpurun_out/step_profile_h3.txt 2>&1
cat gpurun_out/step_profile_h3.txt
timeout 300 python scripts/step_profile.py h3 partial > gpurun_out/step_profile_partial_h3.txt 2>&1
# BASELINE config 4 (scaled inference, 4096 pts, 256 pairs over 8 GPUs = 32 per GPU) and config 5 (kernel sweep)
timeout 600 python bench.py --steps 5 --warmup 3 --workload whole --num-points 4096 --batch 32 --no-cpu-baseline --no-other-workloads > gpurun_out/bench_whole_4096_b32.json 2> gpurun_out/bench_whole_4096_b32.err
cat gpurun_out/bench_whole_4096_b32.json; tail -3 gpurun_out/bench_whole_4096_b32.err
timeout 300 python scripts/knn_bench.py > gpurun_out/knn_bench.txt 2>&1
# CTA-pair / quad GEMM variants vs the single-CTA kernel (bit-identity + timing), small-batch latency eager vs CUDA graph,
# and the default bench with the pair kernel off (A/B record)
timeout -s KILL 150 python scripts/pair_diag.py > gpurun_out/pair_diag.txt 2>&1
timeout -s KILL 200 python scripts/graph_latency.py > gpurun_out/graph_latency.txt 2>&1
VCR_GEMM_PAIR=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-workloads > gpurun_out/bench_default_pair0.json 2> gpurun_out/bench_default_pair0.err
timeout 600 python scripts/kernel_sweep.py > gpurun_out/kernel_sweep_cfg5.txt 2>&1
if [ "$1" != "nonc" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_h3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-workloads > gpurun_out/ncu_launch.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_h3.csv > gpurun_out/launches_h3_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tc|flash_attn|knn_select|edgeconv_dg_tc|attn_colsum|softcorr_tc' -s 60 -c 28 \
    -f -o /tmp/prof_h3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads > gpurun_out/ncu_full.log 2>&1
# the .ncu-rep (60 MB) would push gpurun_out/ past the 64 MiB that travel back: export what the summaries need here
ncu -i /tmp/prof_h3.ncu-rep --page raw --csv > gpurun_out/prof_h3_raw.csv 2>> gpurun_out/ncu_full.log
python scripts/summarize_ncu_full.py gpurun_out/prof_h3_raw.csv > gpurun_out/ncu_full_summary.txt
python scripts/ncu_traffic.py gpurun_out/prof_h3_raw.csv > gpurun_out/ncu_traffic.json
fi
