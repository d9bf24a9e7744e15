#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
for pol in auto 1 0; do
  VCR_GEMM_PAIR=$pol timeout 600 python bench.py --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_pair_$pol.json 2> gpurun_out/bench_pair_$pol.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_pair_$pol.json'))
print("pair=$pol value", round(d['value'],1), "ms", round(d['ms_per_step'],3), d['clocks']['sm_mhz'], "gemm_tc ms", d['kernel_ms_per_step']['vcr_gemm_tc'])
PY
done
timeout 300 python scripts/pair_diag.py 2>&1 | tail -30
