#!/bin/bash
# One GPU-box pass: parity tests, bench (h3 default), per-call profile, ncu launch list, ncu --set full of top kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_h3.json 2> gpurun_out/bench_h3.err
cat gpurun_out/bench_h3.json
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload partial --no-cpu-baseline > gpurun_out/bench_partial_h3.json 2> gpurun_out/bench_partial.err
timeout 300 python scripts/step_profile.py h3 > gpurun_out/step_profile_h3.txt 2>&1
timeout 300 python scripts/step_profile.py h3 partial > gpurun_out/step_profile_h3_partial.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_h3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|flash_attn_tc_kernel|knn_topk_kernel|edgeconv_dg_kernel' -s 40 -c 14 -o gpurun_out/prof_h3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
