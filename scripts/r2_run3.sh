#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -v "parity achieved\|dropin main" gpurun_out/pytest_gpu.log | tail -25
timeout 300 python scripts/step_profile.py h3 partial > gpurun_out/step_profile_partial_h3.txt 2>&1
head -48 gpurun_out/step_profile_partial_h3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'])
print(d['kernel_ms_per_step'])
PY
tail -3 gpurun_out/bench_quick.err
