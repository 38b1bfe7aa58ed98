#!/usr/bin/env python
"""Run flowmap+FTLE on an n x n double-gyre grid (device-resident) -- profiling driver."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.diagnostics import flowmap_ftle_grid_2D
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
x = torch.linspace(0, 2, n, dtype=torch.float64, device="cuda"); y = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
for _ in range(reps):
    fm, ft = flowmap_ftle_grid_2D(f, 0., -10., x, y, p, 2.0 / (n - 1), 1.0 / (n - 1), device_out=True)
torch.cuda.synchronize()
print("ok", float(ft.max()))
