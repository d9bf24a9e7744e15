#!/bin/bash
# One GPU-box pass: parity tests, bench (h3 default), per-call profile.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_h3.json 2> gpurun_out/bench_h3.err
cat gpurun_out/bench_h3.json
timeout 300 python scripts/step_profile.py h3 > gpurun_out/step_profile_h3.txt 2>&1
cat gpurun_out/step_profile_h3.txt
