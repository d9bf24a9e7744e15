#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm or flash or transformer or mha or vcrnet or throughput" | tail -3
timeout -s KILL 200 python scripts/flash_diag.py | sed -n 4,9p
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'], d['gpu_launches'])
print(d['kernel_ms_per_step'])
PY
tail -3 gpurun_out/bench_quick.err
timeout 300 python scripts/step_profile.py 2>&1 | head -24
