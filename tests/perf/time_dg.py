#!/usr/bin/env python
"""Time the double-gyre flow map + FTLE (device-resident, CUDA events) and check parity on config 1.

    [B200CS_LIB=...] python tests/perf/time_dg.py [n=8192] [reps=3]
Prints M points/s (best and median of reps), and for the 401x201 README case the number of
particles whose step sequence differs from the CPU oracle's and max|dx| over the rest."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import oracle
from numbacs_b200 import _lib
from numbacs_b200.diagnostics import flowmap_ftle_grid_2D
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import flowmap_grid_2D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
x = torch.linspace(0, 2, n, dtype=torch.float64, device="cuda")
y = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
ts = []
for r in range(reps + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fm, ft = flowmap_ftle_grid_2D(f, 0., -10., x, y, p, 2.0 / (n - 1), 1.0 / (n - 1), device_out=True)
    e1.record()
    torch.cuda.synchronize()
    if r >= 2:
        ts.append(e0.elapsed_time(e1))
ts = np.array(ts)
print(f"lib={os.path.basename(_lib.LIB_PATH)} n={n}: best {n * n / ts.min() / 1e3:.1f} M pts/s, "
      f"median {n * n / np.median(ts) / 1e3:.1f} M pts/s ({ts.min():.2f} ms)")
xc, yc = np.linspace(0, 2, 401), np.linspace(0, 1, 201)
info = {}
g = flowmap_grid_2D(f, 0., -10., xc, yc, p, info=info)
fo, po, _ = oracle.get_predefined_flow("double_gyre", int_direction=-1.0)
o, _, _, steps_o, stats_o = oracle.flowmap_grid_2D(fo, 0., -10., xc, yc, po, full=True)
same = (info["steps"] == steps_o).all(-1)
d = np.abs(g - o).max(-1)
print(f"  C1 parity: step-sequence mismatches {int((~same).sum())}/{same.size}, "
      f"max|dx| (matching) {d[same].max():.2e}, nfev gpu/oracle {int(info['stats'][0])}/{int(stats_o[0])}")
bad = np.argwhere(~same)
for (i, j) in bad[:8]:
    print(f"    mismatch at x={xc[i]:.4f} y={yc[j]:.4f}: steps gpu {info['steps'][i, j]} oracle {steps_o[i, j]}, "
          f"|dx| = {d[i, j]:.2e}")
print(f"  particles with |dx| > 1e-8: {int((d > 1e-8).sum())}, max|dx| all {d.max():.2e}")
