#!/usr/bin/env python
"""Throughput + parity spot checks for BASELINE configs 1-4 (config 5 is bench.py).

    python tests/perf/bench_configs.py > gpurun_out/configs.json

Each config runs through the public Python API with device-resident tensors, is timed with CUDA
events (best of 3 after a warm-up), and a random subsample of its particles is re-integrated by the
CPU oracle for a parity figure (max |dx| / domain over particles with equal step counts, number of
step-count mismatches)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import oracle as O
from numbacs_b200.flows import (get_predefined_flow, get_interp_arrays_2D, get_flow_2D,
                                get_interp_arrays_scalar, get_callable_scalar)
from numbacs_b200.integration import flowmap_grid_2D, flowmap
from numbacs_b200.diagnostics import ftle_grid_2D, lavd_flowmap_grid_2D

dev = "cuda"


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def spot_parity(f_gpu, f_ora, t0, T, pts, params, L):
    info = {}
    g = flowmap(f_gpu, t0, T, pts, params, info=info)
    o, _, _, steps_o, _ = O.flowmap_pts(f_ora, t0, T, pts, params, full=True)
    d = (np.abs(g - o) / np.asarray(L)).max(axis=-1)
    same = (info["steps"] == steps_o).all(axis=-1)
    return {"sample": len(pts), "step_mismatches": int((~same).sum()),
            "max_rel_dx_matching": float(d[same].max()), "median_rel_dx": float(np.median(d)),
            "p99_rel_dx": float(np.percentile(d, 99))}


res = {}
rng = np.random.default_rng(0)

# ---- C1: double gyre 401 x 201, T = -10
x, y = np.linspace(0, 2, 401), np.linspace(0, 1, 201)
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
fo, po, _ = O.get_predefined_flow("double_gyre", int_direction=-1.0)
xd, yd = torch.tensor(x, device=dev), torch.tensor(y, device=dev)
ms, _ = timed(lambda: ftle_grid_2D(flowmap_grid_2D(f, 0.0, -10.0, xd, yd, p), -10.0, x[1], y[1]))
X, Y = np.meshgrid(x, y, indexing="ij")
pts = np.column_stack((X.ravel(), Y.ravel()))[rng.choice(X.size, 4000, replace=False)]
res["C1 double_gyre 401x201 T=-10"] = {"ms": ms, "Mpts_per_s": X.size / ms / 1e3,
                                       "parity": spot_parity(f, fo, 0.0, -10.0, pts, p, (2.0, 1.0))}

# ---- C2: bickley jet 2001 x 601, T = +6
fb, pb, dom = get_predefined_flow("bickley_jet")
fbo, pbo, _ = O.get_predefined_flow("bickley_jet")
xb, yb = np.linspace(dom[0][0], dom[0][1], 2001), np.linspace(-3, 3, 601)
xbd, ybd = torch.tensor(xb, device=dev), torch.tensor(yb, device=dev)
ms, _ = timed(lambda: ftle_grid_2D(flowmap_grid_2D(fb, 0.0, 6.0, xbd, ybd, pb), 6.0, xb[1] - xb[0], yb[1] - yb[0]))
pts = np.column_stack((rng.uniform(dom[0][0], dom[0][1], 4000), rng.uniform(-3, 3, 4000)))
res["C2 bickley_jet 2001x601 T=+6"] = {"ms": ms, "Mpts_per_s": 2001 * 601 / ms / 1e3,
                                       "parity": spot_parity(fb, fbo, 0.0, 6.0, pts, pb, (dom[0][1], 6.0))}

# ---- C3: MERRA-shaped synthetic field, nt=720 hourly, 576 x 361 lon-lat, FTLE grid 676 x 251, T=-72 h
t = torch.arange(720, dtype=torch.float64, device=dev)
lon = -180.0 + 0.625 * torch.arange(576, dtype=torch.float64, device=dev)
lat = -90.0 + 0.5 * torch.arange(361, dtype=torch.float64, device=dev)
Tm, LO, LA = torch.meshgrid(t, torch.deg2rad(lon), torch.deg2rad(lat), indexing="ij")
U = torch.zeros_like(Tm)
V = torch.zeros_like(Tm)
g = np.random.default_rng(0)
for _ in range(8):
    k, l = int(g.integers(1, 5)), int(g.integers(1, 4))
    ph, om = float(g.uniform(0, 6.28)), float(g.uniform(0.01, 0.05))
    au, av = float(g.uniform(5, 12)), float(g.uniform(3, 8))
    U += au * torch.cos(LA) * torch.sin(k * LO + om * Tm + ph) * torch.cos(l * LA)
    V += av * torch.cos(LA) * torch.cos(k * LO - om * Tm + ph) * torch.sin(2 * l * LA)
del Tm, LO, LA
t0 = time.time()
grid, Cu, Cv = get_interp_arrays_2D(t.cpu().numpy(), lon.cpu().numpy(), lat.cpu().numpy(), U, V)
torch.cuda.synchronize()
t_pref = time.time() - t0
fs = get_flow_2D(grid, Cu, Cv, spherical=1, extrap_mode="linear")
lonf, latf = np.arange(-100, 35 + 0.1, 0.2), np.arange(-5, 45 + 0.1, 0.2)
lond, latd = torch.tensor(lonf, device=dev), torch.tensor(latf, device=dev)
pm = np.array([-1.0])
ms, _ = timed(lambda: ftle_grid_2D(flowmap_grid_2D(fs, 360.0, -72.0, lond, latd, pm), -72.0, 0.2, 0.2))
entry = {"ms": ms, "Mpts_per_s": len(lonf) * len(latf) / ms / 1e3, "grid": [len(lonf), len(latf)],
         "prefilter_s_2x(720x576x361)": t_pref, "coef_GB_on_device": 2 * Cu.numel() * 8 / 1e9}
# oracle spot check needs the coefficients on the host (2 x 1.2 GB)
Cu_h, Cv_h = Cu.cpu().numpy(), Cv.cpu().numpy()
fso = O.get_flow_2D(grid, Cu_h, Cv_h, spherical=1, extrap_mode="linear")
pts = np.column_stack((rng.uniform(-100, 35, 1500), rng.uniform(-5, 45, 1500)))
entry["parity"] = spot_parity(fs, fso, 360.0, -72.0, pts, pm, (360.0, 180.0))
res["C3 MERRA-shaped spline spherical=1 676x251 T=-72h"] = entry
del U, V, Cu, Cv, Cu_h, Cv_h, fso
torch.cuda.empty_cache()

# ---- C4: LAVD, 1024 x 1024 particles, n = 601, QGE-shaped field (257 x 513 x 101), fused path
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
xpd, ypd = torch.tensor(xp, device=dev), torch.tensor(yp, device=dev)
one = np.array([1.0])
ms, out = timed(lambda: lavd_flowmap_grid_2D(fq, 0.5, 0.3, xpd, ypd, one, w, n=601), reps=2)
entry = {"ms": ms, "Mpts_per_s": 1024 * 1024 / ms / 1e3, "n": 601,
         "trajectory_array_avoided_GB": 1024 * 1024 * 601 * 16 / 1e9}
# parity of the fused LAVD against the oracle's two-step path on a 24 x 24 sub-grid
xs, ys = xp[::44], yp[::44]
fqo = O.get_flow_2D(gq, Cuq, Cvq, extrap_mode="linear")
wo = O.get_callable_scalar(gw, Cw, extrap_mode="linear")
Xs, Ys = np.meshgrid(xs, ys, indexing="ij")
got, ts = lavd_flowmap_grid_2D(fq, 0.5, 0.3, xs, ys, one, w, n=601)
fmno, tso = O.flowmap_n_grid_2D(fqo, 0.5, 0.3, xs, ys, one, n=601)
ref = O.lavd_grid_2D(fmno, tso, 0.3, wo, Xs.ravel(), Ys.ravel())
entry["parity"] = {"sample": int(Xs.size), "lavd_rel_L2": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
                   "lavd_max_abs": float(np.abs(got - ref).max()), "lavd_scale": float(np.abs(ref).max())}
res["C4 LAVD fused 1024x1024 n=601 QGE-shaped"] = entry

print(json.dumps(res, indent=1))
