#!/bin/bash
# round 2, GPU pass 1: full GPU suite (incl. headline parity, DataParallel on 2 devices, main.py drop-in), step profile
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python scripts/step_profile.py h3 partial > gpurun_out/step_profile_partial_h3.txt 2>&1
tail -50 gpurun_out/step_profile_partial_h3.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
