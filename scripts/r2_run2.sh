#!/bin/bash
# round 2, GPU pass 2: full GPU suite, step profile (hoisted loop), default bench + reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/step_profile.py h3 partial > gpurun_out/step_profile_partial_h3.txt 2>&1
head -45 gpurun_out/step_profile_partial_h3.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_reference.json; tail -3 gpurun_out/bench_reference.err
