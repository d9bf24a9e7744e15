set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/dev_pytest_gpu.log 2>&1; tail -8 gpurun_out/dev_pytest_gpu.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/dev_bench_partial.json 2>gpurun_out/dev_bench_partial.err
python - <<PY
import json
d=json.load(open("gpurun_out/dev_bench_partial.json"))
print("partial", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "whole", {k[:20]:round(v["value"],1) for k,v in d.get("other_workloads",{}).items()})
k=d["kernel_ms_per_step"]
for n,v in sorted(k.items(), key=lambda kv:-kv[1])[:14]: print(f"{v:7.3f} {n}")
PY
