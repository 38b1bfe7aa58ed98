#!/usr/bin/env python
"""Profiling / timing driver of BASELINE config 4: fused flow map + LAVD (lavd_flowmap_kernel), 1024 x 1024
particles, n = 601 output times, QGE-shaped cubic-spline velocity and vorticity (257 x 513 x 101).

    python tools/prof_lavd.py [reps=3]         -> one JSON line (ms, nfev / particle, parity vs the oracle on a sub-grid)
    ncu --set full --clock-control none --import-source on -k regex:lavd_flowmap_kernel -s 1 -c 1 \
        -o gpurun_out/lavd python tools/prof_lavd.py 2"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from numbacs_b200 import _lib
from numbacs_b200.diagnostics import lavd_flowmap_grid_2D
from numbacs_b200.flows import get_callable_scalar, get_flow_2D, get_interp_arrays_2D, get_interp_arrays_scalar

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = np.random.default_rng(0)
xq, yq, tq = np.linspace(0, 1, 257), np.linspace(0, 2, 513), np.linspace(0, 1, 101)
Tq, Xq, Yq = np.meshgrid(tq, xq, yq, indexing="ij")
psi = np.zeros_like(Tq)
for _ in range(6):
    k, l = int(g.integers(1, 4)), int(g.integers(1, 5))
    amp, om, ph = float(g.uniform(0.02, 0.06)), float(g.uniform(1, 6)), float(g.uniform(0, 6.28))
    psi += amp * np.sin(k * np.pi * Xq) * np.sin(l * np.pi * Yq / 2) * np.cos(om * Tq + ph)
dxq, dyq = xq[1] - xq[0], yq[1] - yq[0]
Uq = -np.gradient(psi, dyq, axis=2)
Vq = np.gradient(psi, dxq, axis=1)
vort = np.gradient(Vq, dxq, axis=1) - np.gradient(Uq, dyq, axis=2)
gq, Cuq, Cvq = get_interp_arrays_2D(tq, xq, yq, Uq, Vq)
fq = get_flow_2D(gq, Cuq, Cvq, extrap_mode="linear")
gw, Cw = get_interp_arrays_scalar(tq, xq, yq, vort)
w = get_callable_scalar(gw, Cw, extrap_mode="linear")
xp, yp = np.linspace(0.02, 0.98, 1024), np.linspace(0.02, 1.98, 1024)
xpd, ypd = torch.tensor(xp, device="cuda"), torch.tensor(yp, device="cuda")
one = np.array([1.0])
ts, info = [], {}
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    info = {}
    e0.record()
    out = lavd_flowmap_grid_2D(fq, 0.5, 0.3, xpd, ypd, one, w, n=601, info=info)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
lavd = out[0] if isinstance(out, tuple) else out
res = {"lib": os.path.basename(_lib.LIB_PATH), "ms": min(ts[1:] or ts), "particles": 1024 * 1024, "n": 601,
       "checksum": float(torch.as_tensor(lavd).double().sum())}
if "stats" in info:
    st = np.asarray(info["stats"].cpu() if hasattr(info["stats"], "cpu") else info["stats"], dtype=np.float64)
    res["nfev_per_particle"] = st[0] / (1024 * 1024)
    res["attempts_per_particle"] = (st[1] + st[2]) / (1024 * 1024)
print(json.dumps(res))
