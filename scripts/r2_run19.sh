#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -q -x -k "select or vcp or vcrnet or headline or colsum or flash" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'], d['gpu_launches'])
print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.08})
PY
tail -3 gpurun_out/bench_quick.err
