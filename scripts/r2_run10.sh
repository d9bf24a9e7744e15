#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print("value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'], d['gpu_launches'])
print("ref on gpu:", d.get('reference_on_this_gpu'))
print("cpu:", d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
print({k:(round(v['value']),v.get('scaling')) for k,v in d['other_workloads'].items()})
PY
