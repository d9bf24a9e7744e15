"""Summarise an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv > raw.csv; python scripts/summarize_ncu_full.py raw.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct"]
ix = [hdr.index(w) for w in want if w in hdr]
print("# " + ", ".join(hdr[i] + (" [" + units[i] + "]" if units[i] else "") for i in ix))
for r in rows[2:]:
    print(", ".join(r[i][:48] for i in ix))
