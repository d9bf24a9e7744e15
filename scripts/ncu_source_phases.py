"""Aggregate an `ncu --page source --csv` export (SASS view) by barrier-delimited phase and by opcode."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
nq = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # first kernel section only
    if len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
iS, iN, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN]) for r in data); tots = sum(int(r[iSm]) for r in data)
print("total warp-instr", tot, "samples", tots, "per unit:", tot / nq)
phase = 0
ph = collections.defaultdict(lambda: [0, 0, collections.Counter()])
ops = collections.defaultdict(lambda: [0, 0])
for r in data:
    src = r[iS].strip()
    parts = src.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    n = int(r[iN]); sm = int(r[iSm])
    ph[phase][0] += n; ph[phase][1] += sm
    for i in stall: ph[phase][2][hdr[i]] += int(r[i])
    ops[op][0] += n; ops[op][1] += sm
    if op == 'BAR': phase += 1
for k, v in ph.items():
    print("phase", k, "instr %.1f%% samples %.1f%%" % (100 * v[0] / tot, 100 * v[1] / tots), v[2].most_common(5))
for k, v in sorted(ops.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k:10s} instr {100*v[0]/tot:5.1f}%  samples {100*v[1]/tots:5.1f}%")
