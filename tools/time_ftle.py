#!/usr/bin/env python
"""Time ftle_grid_2D alone on an n x n double-gyre flow map (device-resident, CUDA events).
    [B200CS_LIB=...] python tools/time_ftle.py [n=16384] [reps=10]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numbacs_b200 import _lib
from numbacs_b200.diagnostics import ftle_grid_2D
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import flowmap_grid_2D
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
x = torch.linspace(0, 2, n, dtype=torch.float64, device="cuda"); y = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
fm = flowmap_grid_2D(f, 0., -10., x, y, p, device_out=True)
ts = []
for r in range(reps + 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ft = ftle_grid_2D(fm, -10., 2.0 / (n - 1), 1.0 / (n - 1)); e1.record(); torch.cuda.synchronize()
    if r >= 2: ts.append(e0.elapsed_time(e1))
t = min(ts)
print(f"lib={os.path.basename(_lib.LIB_PATH)} n={n}: ftle {t:.3f} ms = {24.0 * n * n / t / 1e6:.0f} GB/s algorithmic "
      f"(median {np.median(ts):.3f} ms), checksum {float(ft.sum()):.10f}")
