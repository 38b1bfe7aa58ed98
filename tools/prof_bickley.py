#!/usr/bin/env python
"""Profiling driver for the Bickley-jet flow-map kernel (BASELINE config 2: 2001 x 601, T = +6),
device-resident.  Under ncu:
    ncu --set full --clock-control none --import-source on -k regex:flowmap_queue_kernel -s 1 -c 1 \
        -o gpurun_out/bickley python tools/prof_bickley.py [scale=1] [reps=2]
`scale` multiplies both grid dimensions (scale=3 fills the machine: 10.8 M particles)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import flowmap_grid_2D

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
f, p, dom = get_predefined_flow("bickley_jet")
x = torch.linspace(dom[0][0], dom[0][1], 2001 * scale, dtype=torch.float64, device="cuda")
y = torch.linspace(-3, 3, 601 * scale, dtype=torch.float64, device="cuda")
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    info = {}
    e0.record()
    fm = flowmap_grid_2D(f, 0.0, 6.0, x, y, p, device_out=True, info=info)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
n = x.numel() * y.numel()
st = np.asarray(info["stats"].cpu() if hasattr(info["stats"], "cpu") else info["stats"], dtype=np.float64)
print(f"bickley {x.numel()} x {y.numel()}: {min(ts):.3f} ms = {n / min(ts) / 1e3:.1f} M points/s; "
      f"attempts/particle {(st[1] + st[2]) / n:.2f}, nfev/particle {st[0] / n:.1f}")
