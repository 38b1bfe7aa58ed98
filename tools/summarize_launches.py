#!/usr/bin/env python
"""Per-kernel totals from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: tools/summarize_launches.py gpurun_out/launches.csv"""
import csv, sys
from collections import defaultdict
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
iK, iV, iM = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = defaultdict(list)
for r in rows[1:]:
    if len(r) > iV and r[iM] == "gpu__time_duration.sum":
        agg[r[iK][:78]].append(float(r[iV].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"# {sys.argv[1]}: {sum(len(v) for v in agg.values())} launches, {tot / 1e6:.2f} ms of kernel time "
      "(ncu: cold-cache, serialised -- compare shares, not absolutes)")
print(f"{'kernel':80s} {'n':>4s} {'total ms':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:80s} {len(v):4d} {sum(v) / 1e6:10.3f} {100 * sum(v) / tot:6.2f}%")
