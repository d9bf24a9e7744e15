#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for n in 1 $N; do
  if [ "$n" -eq 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_scale_1gpu.json 2> gpurun_out/bench_scale_1gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_scale_${n}gpu.json 2> gpurun_out/bench_scale_${n}gpu.err
  fi
  tail -2 gpurun_out/bench_scale_${n}gpu.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_scale_${n}gpu.json'))
print("N", d['n_gpus'], "value", round(d['value'],1), "ms", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value'],1), d['clocks'], "launches", d['gpu_launches'])
print({k:(round(v['value']),v.get('scaling')) for k,v in d['other_workloads'].items()})
print({k:round(v['value']) for k,v in d.items() if k.startswith('variant')}, d.get('latency_batch1'))
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_reference_${N}gpu.json 2> gpurun_out/bench_reference_${N}gpu.err
tail -c 400 gpurun_out/bench_reference_${N}gpu.json
