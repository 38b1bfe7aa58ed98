#!/usr/bin/env python
"""Dynamic opcode mix + top stall sites from an `ncu --page source --csv` dump.
usage: ncu -i rep --page source --csv > src.csv; tools/ncu_opmix.py src.csv [top]"""
import csv, sys, re
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp = Counter(), Counter()
tot = 0
lines = []
for r in rows[2:]:
    if len(r) <= iE or not r[iE]:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    if not m:
        continue
    n = int(r[iE]); s = int(r[iSamp] or 0)
    ops[m.group(2)] += n; samp[m.group(2)] += s; tot += n
    lines.append((s, n, r[0], r[iS].strip()))
print("total warp instructions executed:", tot)
fp64 = sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
print("FP64-pipe (DFMA+DMUL+DADD+DSETP):", sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP")), "share %.1f%%" % (100.0 * sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP")) / tot))
for k, v in ops.most_common(top):
    print(f"{k:10s} {v:14d} {100.0*v/tot:5.1f}%   samples {samp[k]}")
print("--- top stall sites")
for s, n, a, src in sorted(lines, reverse=True)[:top]:
    print(s, n, a, src[:90])
