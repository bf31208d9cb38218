"""Basic-block view of an ncu source-page CSV: runs of SASS lines with equal executed count, with samples and opcode mix.
usage: python tools/ncu_blocks.py x_source.csv [min_total_M]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
hdr = rows[1]; isrc = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
R = rows[2:]
tot_s = sum(int(r[isamp] or 0) for r in R if len(r) > isamp and (r[isamp] or "0").isdigit())
prev = None; start = 0; acc = 0; ops = {}
def flush(end):
    if prev is None: return
    n = end - start
    if prev * n / 1e6 < thr and acc < 0.01 * tot_s: return
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:7]
    print(f"{start:5d}-{end-1:5d} n={n:4d} exec/inst={prev:10d} total={prev*n/1e6:8.1f}M samples={acc:6d} ({100*acc/tot_s:4.1f}%) {top}")
for n, r in enumerate(R):
    try: e = int(r[iex] or 0); s = int(r[isamp] or 0)
    except (ValueError, IndexError): continue
    src = r[isrc]
    if e != prev or 'BAR' in src:
        flush(n); prev = e; start = n; acc = 0; ops = {}
    acc += s
    op = (src.split()[1] if src.startswith('@') else src.split()[0]).split('.')[0]
    if 'BAR' in src: op = src[:30]
    ops[op] = ops.get(op, 0) + 1
flush(len(R))
