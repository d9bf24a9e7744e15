#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flash or colsum" | tail -5
timeout -s KILL 300 python scripts/flash_diag.py | tee gpurun_out/flash_diag.txt
for fw in 3 5; do
VCR_FLASH_WARPS=$fw timeout 600 python bench.py --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_quick_fw$fw.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_quick_fw$fw.json'))
print("fw$fw value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'], d['gpu_launches'])
print({k:v for k,v in d['kernel_ms_per_step'].items() if v>0.1})
PY
tail -3 gpurun_out/bench_quick.err
done
