#!/usr/bin/env python
"""Static opcode mix of the step-attempt loop of a flow-map kernel (no GPU needed).

    tools/sass_loop_mix.py <object-or-so> <kernel-name-substring>

The attempt loop is straight-line code (stages fully unrolled), so the static mix of the largest
backward-branch span is a good proxy for the dynamic per-attempt mix: FP64-pipe instructions
(DFMA/DMUL/DADD/DSETP, two pipe cycles each) against everything else (one issue slot each)."""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
funcs = [l.split()[2] for l in subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
         if "Function :" in l and pat in l]
fn = funcs[0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
best = (0, 0, 0)
for a, op, rest in ins:
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", rest)
        if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > best[0]:
            best = (a - int(m.group(1), 16), int(m.group(1), 16), a)
_, lo, hi = best
loop = [(a, op) for a, op, _ in ins if lo <= a <= hi]
c = Counter(op.split(".")[0] for _, op in loop)
fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"{fn[-60:]}: {len(ins)} instructions, loop 0x{lo:x}-0x{hi:x} = {len(loop)}")
print(f"  FP64 {fp64}  other {len(loop) - fp64}  | " + "  ".join(f"{k} {v}" for k, v in c.most_common(24)))
