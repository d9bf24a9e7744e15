set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn" > gpurun_out/dev_knn_pytest.log 2>&1; tail -15 gpurun_out/dev_knn_pytest.log
timeout 200 python scripts/knn_bench.py > gpurun_out/dev_knn_bench.txt 2>&1; cat gpurun_out/dev_knn_bench.txt
for mb in 3072 96 64 40; do
VCR_STAT_CHUNK_MB=$mb timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-other-workloads > gpurun_out/dev_bench_chunk$mb.json 2>gpurun_out/dev_bench_chunk$mb.err
python - <<PY
import json
d=json.load(open("gpurun_out/dev_bench_chunk$mb.json"))
print("chunk_mb=$mb", round(d["value"],1), "pairs/s", {k:v for k,v in d["kernel_ms_per_step"].items() if k in ("vcr_gemm_tc","vcr_softmax_colsum")})
PY
done
