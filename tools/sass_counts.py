#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libcngp.so (cuobjdump -sass): which pipes the shipped code uses.
    python tools/sass_counts.py > profiles/sass_counts_r02.txt
DMMA = FP64 tensor core (mma.sync.m8n8k4.f64), HMMA.1688.F32.TF32 = TF32 tensor core (FP32 mode), UBLKCP = bulk TMA copy
(cp.async.bulk), SYNCS = mbarrier operations, LDS/STS = shared memory, DFMA/DMUL/DADD = scalar FP64."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "corenav_gp_b200", "libcngp.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = {"DMMA": r"\bDMMA", "HMMA(TF32)": r"\bHMMA\.\d+\.F32\.TF32", "UBLKCP": r"\bUBLKCP", "SYNCS": r"\bSYNCS",
       "UTCMMA/LDTM": r"\b(UTC\w*MMA|LDTM)", "LDS": r"\bLDS", "STS": r"\bSTS", "DFMA": r"\bDFMA", "DMUL": r"\bDMUL",
       "DADD": r"\bDADD", "LDL/STL": r"\b(LDL|STL)", "BAR": r"\bBAR\."}
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pat.items():
            if re.search(p, line):
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
print("# " + "  ".join(f"{k:>11s}" for k in pat) + "  kernel")
tot = collections.Counter()
for (mangled, c), name in zip(counts.items(), names):
    tot.update(c)
    short = re.sub(r"\(cngp::\w+\)|\(\(anonymous namespace\)::\w+\)", "", name).replace("cngp::", "").replace("(anonymous namespace)::", "")
    print("  " + "  ".join(f"{c[k]:11d}" for k in pat) + "  " + short[:110])
print("  " + "  ".join(f"{tot[k]:11d}" for k in pat) + "  TOTAL")
