#!/bin/bash
exec 2>&1
# default bench as the driver runs it (timed), reference arm, 2-GPU torchrun
mkdir -p gpurun_out
time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -4 gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.json; echo
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_scale_${N}gpu.json 2> gpurun_out/bench_scale_${N}gpu.err
tail -4 gpurun_out/bench_scale_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_scale_${N}gpu.json'))
print("N", d['n_gpus'], "value", d['value'], "ms", d['ms_per_step'], "e2e", d['e2e']['value'], d['clocks'], "launches", d['gpu_launches'])
print({k:(round(v['value']),v.get('scaling')) for k,v in d['other_workloads'].items()})
print({k:round(v['value']) for k,v in d.items() if k.startswith('variant')})
PY
fi
