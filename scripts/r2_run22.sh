#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "parity achieved\|dropin main" | tail -4 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'], d['gpu_launches'])
print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.04})
PY
tail -2 gpurun_out/bench_quick.err
