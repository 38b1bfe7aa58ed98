#!/usr/bin/env python
"""Config-3-shaped spline flow at a particle count that fills the GPU: timing + gather bandwidth.

    python tools/prof_spline.py [step_deg=0.05] [reps=3]        -> one JSON line
    ncu --set full -k regex:flowmap_kernel -c 1 ... python tools/prof_spline.py 0.05 1

Synthetic MERRA-shaped velocity (nt = 720 hourly, 576 x 361 lon-lat), cubic-spline flow with
spherical = 1, particles on lon [-100, 35] x lat [-5, 45] at `step_deg`, t0 = 360, T = -72 h.
Reports M points/s and the ALGORITHMIC gather traffic of the RHS: 64 taps x 16 B (u, v interleaved)
= 1 KiB per RHS evaluation, times the kernel's own count of RHS evaluations."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from numbacs_b200.flows import get_flow_2D, get_interp_arrays_2D
from numbacs_b200.integration import flowmap_grid_2D

step = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
t = torch.arange(720, dtype=torch.float64, device=dev)
lon = -180.0 + 0.625 * torch.arange(576, dtype=torch.float64, device=dev)
lat = -90.0 + 0.5 * torch.arange(361, dtype=torch.float64, device=dev)
Tm, LO, LA = torch.meshgrid(t, torch.deg2rad(lon), torch.deg2rad(lat), indexing="ij")
U, V = torch.zeros_like(Tm), torch.zeros_like(Tm)
g = np.random.default_rng(0)
for _ in range(8):
    k, l = int(g.integers(1, 5)), int(g.integers(1, 4))
    ph, om = float(g.uniform(0, 6.28)), float(g.uniform(0.01, 0.05))
    au, av = float(g.uniform(5, 12)), float(g.uniform(3, 8))
    U += au * torch.cos(LA) * torch.sin(k * LO + om * Tm + ph) * torch.cos(l * LA)
    V += av * torch.cos(LA) * torch.cos(k * LO - om * Tm + ph) * torch.sin(2 * l * LA)
del Tm, LO, LA
grid, Cu, Cv = get_interp_arrays_2D(t.cpu().numpy(), lon.cpu().numpy(), lat.cpu().numpy(), U, V)
fs = get_flow_2D(grid, Cu, Cv, spherical=1, extrap_mode="linear")
del U, V, Cu, Cv
torch.cuda.empty_cache()
lonf, latf = np.arange(-100, 35 + step / 2, step), np.arange(-5, 45 + step / 2, step)
lond, latd = torch.tensor(lonf, device=dev), torch.tensor(latf, device=dev)
pm = np.array([-1.0])
best, stats = float("inf"), None
for r in range(reps + (1 if reps > 1 else 0)):
    info = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fm = flowmap_grid_2D(fs, 360.0, -72.0, lond, latd, pm, device_out=True, info=info)
    e1.record()
    torch.cuda.synchronize()
    if r > 0 or reps == 1:
        best = min(best, e0.elapsed_time(e1))
    stats = info["stats"].cpu().numpy()
npts = len(lonf) * len(latf)
nfev = float(stats[0])
print(json.dumps({"grid": [len(lonf), len(latf)], "particles": npts, "ms": best,
                  "Mpts_per_s": npts / best / 1e3, "nfev_per_particle": nfev / npts,
                  "attempts_per_particle": float(stats[1] + stats[2]) / npts,
                  "gather_bytes_per_rhs": 1024, "gather_GB": nfev * 1024 / 1e9,
                  "gather_GBps_algorithmic": nfev * 1024 / best / 1e6,
                  "coef_GB": 2 * 722 * 578 * 363 * 8 / 1e9}))
