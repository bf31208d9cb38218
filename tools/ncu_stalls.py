"""Stall-reason totals per opcode class from an ncu source-page CSV.
usage: ncu -i rep --page source --csv > x.csv ; python tools/ncu_stalls.py x.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc = hdr.index("Source")
cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); byop = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    try:
        src = r[isrc]
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
        for i, h in cols:
            v = int(r[i] or 0)
            tot[h] += v; byop[op][h] += v
    except (ValueError, IndexError):
        pass
s = sum(tot.values())
print("stall totals:", {k: f"{100*v/s:.1f}%" for k, v in tot.most_common(10)})
for op, c in sorted(byop.items(), key=lambda kv: -sum(kv[1].values()))[:10]:
    t = sum(c.values())
    print(f"{op:8s} {100*t/s:5.1f}%  ", {k[6:]: f"{100*v/t:.0f}%" for k, v in c.most_common(4)})
