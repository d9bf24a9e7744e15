"""From an `ncu --set full` raw CSV: per kernel family, launches, mean duration, mean DRAM traffic (read+write bytes per
launch) and mean tensor-pipe activity.  Output JSON is committed under profiles/ and read by bench.py (`roofline.traffic`).
usage: python scripts/ncu_traffic.py raw.csv [raw2.csv ...] > profiles/rNN_ncu_traffic.json"""
import csv, json, re, sys, collections
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
FAMILY = [("gemm_tc_kernel", "vcr_gemm_tc"), ("flash_attn_ts_kernel", "vcr_flash_attn_tc"), ("flash_attn_tc_kernel", "vcr_flash_attn_tc"), ("knn_select_kernel<16", "vcr_knn_topk[D=64]"),
          ("knn_select_kernel<(int)16", "vcr_knn_topk[D=64]"), ("knn3_kernel", "vcr_knn_topk[D=3]"),
          ("knn_select_kernel", "vcr_knn_topk[D=3]"), ("layernorm_operand_kernel", "vcr_layernorm_operand"),
          ("knn_tc2_kernel", "vcr_knn_topk_tc"), ("edgeconv_dg_tc_kernel", "vcr_edgeconv_dg_tc"),
          ("attn_colsum_tc_kernel", "vcr_attn_colsum_tc"), ("softcorr_tc_kernel", "vcr_softcorr_tc"), ("knn_tc_kernel", "vcr_knn_topk_tc")]
agg = collections.OrderedDict()
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        fam = next((f for pat, f in FAMILY if pat in name), None)
        if fam is None:
            continue
        def val(metric, table):
            i = col[metric]
            return float(r[i].replace(",", "")) * table.get(units[i], 1.0)
        d = agg.setdefault(fam, {"launches": 0, "us": 0.0, "dram_bytes": 0.0, "tensor_pct": 0.0, "fma_pct": 0.0, "source": []})
        d["launches"] += 1
        d["us"] += val("gpu__time_duration.sum", TIME)
        d["dram_bytes"] += val("dram__bytes_read.sum", UNIT) + val("dram__bytes_write.sum", UNIT)
        d["tensor_pct"] += float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
        d["fma_pct"] += float(r[col["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]])
        if path not in d["source"]:
            d["source"].append(path)
out = {}
for fam, d in agg.items():
    n = d["launches"]
    out[fam] = {"launches_captured": n, "mean_us_under_ncu": round(d["us"] / n, 2),
                "mean_dram_bytes_per_launch": round(d["dram_bytes"] / n),
                "mean_tensor_pipe_active_pct": round(d["tensor_pct"] / n, 2),
                "mean_fma_pipe_active_pct": round(d["fma_pct"] / n, 2), "source": d["source"]}
print(json.dumps(out, indent=1))
