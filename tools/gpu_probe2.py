#!/usr/bin/env python
"""RHS-level parity probe (scratch tool)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
from numbacs_b200 import _lib
from numbacs_b200.flows import get_predefined_flow, get_interp_arrays_2D, get_flow_2D
from numbacs_b200.integration import flowmap_grid_2D

def gpu_rhs(h, t, y, p):
    t = np.ascontiguousarray(t); y = np.ascontiguousarray(y); p = np.ascontiguousarray(p, dtype=np.float64)
    dy = np.empty_like(y)
    _lib.check(_lib.load().b200cs_flow_rhs(h, t.ctypes.data, y.ctypes.data, len(t), p.ctypes.data, len(p), dy.ctypes.data, None))
    return dy

rng = np.random.default_rng(0)
for name, lo, hi in (("double_gyre", (0, 0), (2, 1)), ("bickley_jet", (0, -3), (20, 3)), ("abc", (0, 0, 0), (6.28, 6.28, 6.28))):
    h, p, _ = get_predefined_flow(name); ho, po, _ = O.get_predefined_flow(name)
    n = 20000
    y = rng.uniform(lo, hi, size=(n, len(lo))); t = rng.uniform(0, 10, size=n)
    g = gpu_rhs(h, t, y, p)
    o = np.array([ho.rhs(t[i], y[i], po) for i in range(n)])
    scale = np.abs(o).max(axis=0)
    d = np.abs(g - o)
    print(name, "max abs diff per comp", d.max(axis=0), "scale", scale, "rel", d.max(axis=0) / scale, flush=True)
    k = np.unravel_index(np.argmax(d / scale), d.shape)
    print("  worst at", t[k[0]], y[k[0]], g[k[0]], o[k[0]])

tt = np.linspace(0, 10, 21); xs = np.linspace(0, 2, 41); ys = np.linspace(0, 1, 31)
Tm, Xm, Ym = np.meshgrid(tt, xs, ys, indexing="ij")
a = 0.25 * np.sin(0.2 * np.pi * Tm); b = 1 - 2 * a; ff = a * Xm ** 2 + b * Xm
U = -np.pi * 0.1 * np.sin(np.pi * ff) * np.cos(np.pi * Ym); V = np.pi * 0.1 * np.cos(np.pi * ff) * np.sin(np.pi * Ym) * (2 * a * Xm + b)
grid, Cu, Cv = get_interp_arrays_2D(tt, xs, ys, U, V)
fs = get_flow_2D(grid, Cu, Cv); fso = O.get_flow_2D(grid, Cu, Cv)
n = 20000
y = rng.uniform((0.01, 0.01), (1.99, 0.99), size=(n, 2)); t = rng.uniform(0, 10, size=n)
g = gpu_rhs(fs, t, y, np.array([1.0])); o = np.array([fso.rhs(t[i], y[i], np.array([1.0])) for i in range(n)])
d = np.abs(g - o); print("spline rhs max abs diff", d.max(axis=0), "scale", np.abs(o).max(axis=0), flush=True)

# DG analytic on the spline probe's grid/time for comparison of noise amplification
xg = np.linspace(0.05, 1.95, 101); yg = np.linspace(0.05, 0.95, 51)
h, p, _ = get_predefined_flow("double_gyre"); ho, po, _ = O.get_predefined_flow("double_gyre")
info = {}
fm = flowmap_grid_2D(h, 0., 8., xg, yg, p, info=info)
fmo, _, sto, stepso, statso = O.flowmap_grid_2D(ho, 0., 8., xg, yg, po, full=True)
dd = np.abs(fm - fmo).max(axis=-1); same = (info["steps"] == stepso).all(axis=-1)
print("DG analytic same grid: mismatches", int((~same).sum()), "maxdiff", dd[same].max(), "steps mean", stepso.mean(axis=(0, 1)))
info = {}
fm = flowmap_grid_2D(fs, 0., 8., xg, yg, np.array([1.0]), info=info)
fmo, _, sto, stepso, statso = O.flowmap_grid_2D(fso, 0., 8., xg, yg, np.array([1.0]), full=True)
dd = np.abs(fm - fmo).max(axis=-1); same = (info["steps"] == stepso).all(axis=-1)
print("spline: mismatches", int((~same).sum()), "maxdiff", dd[same].max(), "p99", np.percentile(dd, 99), "median", np.median(dd), "steps mean", stepso.mean(axis=(0, 1)))
# bickley detail
hb, pb, dom = get_predefined_flow("bickley_jet"); hbo, pbo, _ = O.get_predefined_flow("bickley_jet")
xb = np.linspace(dom[0][0], dom[0][1], 401); yb = np.linspace(-3, 3, 121)
for T in (1.0, 3.0, 6.0):
    info = {}
    fm = flowmap_grid_2D(hb, 0., T, xb, yb, pb, info=info)
    fmo, _, sto, stepso, statso = O.flowmap_grid_2D(hbo, 0., T, xb, yb, pbo, full=True)
    dd = np.abs(fm - fmo).max(axis=-1); same = (info["steps"] == stepso).all(axis=-1)
    print("bickley T", T, "mismatches", int((~same).sum()), "maxdiff match", dd[same].max(), "p99", np.percentile(dd, 99), "median", np.median(dd), "steps", stepso.mean(axis=(0, 1)), "status!=1", int((info["status"] != 1).sum()), int((sto != 1).sum()))
