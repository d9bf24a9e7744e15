"""Per-kernel SASS-level summary of an `ncu --set full --import-source on` report: instruction mix, warp-stall reasons and
barrier-delimited phases (share of executed warp-instructions vs share of stall samples).
usage: python scripts/ncu_source_mix.py report.ncu-rep id[,id...] > profiles/rNN_ncu_source_mix.txt"""
import collections, csv, io, subprocess, sys

rep, ids = sys.argv[1], [int(x) for x in sys.argv[2].split(",")]
for kid in ids:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(kid), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = rows[1]
    iN, iS, iSm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    seen, data = set(), []
    for r in rows[2:]:
        if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0]); data.append(r)
    tot = sum(int(r[iN]) for r in data); ts = max(1, sum(int(r[iSm]) for r in data))
    print(f"=== launch {kid}: {name[:110]}")
    print(f"    SASS instructions {len(data)}, executed warp-instructions {tot}, stall samples {ts}")
    ops, sr = collections.Counter(), collections.Counter()
    phase, ph = 0, collections.defaultdict(lambda: [0, 0])
    for r in data:
        parts = r[iS].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        n, sm = int(r[iN]), int(r[iSm])
        ops[op] += n
        ph[phase][0] += n; ph[phase][1] += sm
        for i in stall:
            sr[hdr[i]] += int(r[i])
        if op == "BAR":
            phase += 1
    print("    instruction mix: " + ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in ops.most_common(12)))
    sts = max(1, sum(sr.values()))
    print("    stall reasons:   " + ", ".join(f"{o[6:]} {100*n/sts:.1f}%" for o, n in sr.most_common(7)))
    big = [(k, v) for k, v in ph.items() if v[0] / max(tot, 1) > 0.02 or v[1] / ts > 0.02]
    if len(big) > 1:
        print("    phases (between BAR.SYNC): " + "; ".join(f"#{k}: instr {100*v[0]/tot:.0f}% / samples {100*v[1]/ts:.0f}%" for k, v in big))
