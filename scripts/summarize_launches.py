"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: kernel, launches, total us, share."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("<unnamed>::", "")
    d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += v
tot = sum(d[1] for d in agg.values())
print(f"# {sys.argv[1]}: {sum(d[0] for d in agg.values())} launches, {tot:.1f} us total")
print("# kernel, launches, total_us, share")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k}, {n}, {us:.1f}, {us/tot:.3f}")
