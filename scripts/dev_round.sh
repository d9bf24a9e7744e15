set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "attn_colsum or partial or transformer" > gpurun_out/dev_cs_pytest.log 2>&1; tail -15 gpurun_out/dev_cs_pytest.log
for f in 1 0; do
VCR_FUSED_KEY_STAT=$f timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-workloads > gpurun_out/dev_bench_partial_ks$f.json 2>gpurun_out/dev_bench_partial_ks$f.err
python - <<PY
import json
d=json.load(open("gpurun_out/dev_bench_partial_ks$f.json"))
print("fused_key_stat=$f", round(d["value"],1), "pairs/s", d["ms_per_step"], {k:v for k,v in d["kernel_ms_per_step"].items() if "colsum" in k or "gemm_tc" in k})
PY
done
