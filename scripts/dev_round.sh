set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 100 python scripts/step_profile.py h3 partial > gpurun_out/dev_step_profile_partial.txt 2>&1; grep -E "knn|precision" gpurun_out/dev_step_profile_partial.txt
timeout -s KILL 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-workloads 2>/dev/null | cut -c1-160
