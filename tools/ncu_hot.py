"""Top SASS instructions by executed count / stall samples from an ncu source-page CSV.
usage: ncu -i rep --page source --csv > x.csv ; python tools/ncu_hot.py x.csv [n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
data = []
for r in rows[2:]:
    try:
        data.append((int(r[iex] or 0), int(r[isamp] or 0), r[isrc]))
    except (ValueError, IndexError):
        pass
tot_ex = sum(d[0] for d in data); tot_s = sum(d[1] for d in data)
print("total executed", tot_ex, "samples", tot_s, "sass lines", len(data))
byop = collections.Counter(); sop = collections.Counter()
for ex, s, src in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    byop[op] += ex; sop[op] += s
print("by opcode (executed %, samples %):")
for op, ex in byop.most_common(18):
    print(f"  {op:10s} {100*ex/tot_ex:5.1f}%  {100*sop[op]/max(1,tot_s):5.1f}%")
print("top by samples:")
for i, (ex, s, src) in enumerate(sorted(data, key=lambda d: -d[1])[:n]):
    print(f"  {100*s/max(1,tot_s):5.2f}%  ex={ex:10d}  {src[:110]}")
