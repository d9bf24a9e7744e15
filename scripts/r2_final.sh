#!/bin/bash
# round 2 final evidence pass (1 GPU): full GPU suite, default bench, reference arm, ncu launch list + --set full, kernel sweep
exec 2>&1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "parity achieved\|dropin main" gpurun_out/pytest_gpu.log | tail -12
timeout 300 python scripts/step_profile.py h3 partial > gpurun_out/step_profile_partial_h3.txt 2>&1
timeout 300 python scripts/step_profile.py h3 > gpurun_out/step_profile_h3.txt 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 300 gpurun_out/bench_reference.json; echo
bash scripts/r2_profile.sh
