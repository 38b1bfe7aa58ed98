#!/usr/bin/env python
"""First-contact GPU probe: parity numbers against the CPU oracle + rough timings (scratch tool)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
from numbacs_b200 import _lib
from numbacs_b200.flows import (get_predefined_flow, get_interp_arrays_2D, get_flow_2D,
                                get_callable_scalar, get_callable_scalar_linear)
from numbacs_b200.integration import flowmap_grid_2D, flowmap_n_grid_2D, flowmap, flowmap_n
from numbacs_b200.diagnostics import ftle_grid_2D, lavd_grid_2D, flowmap_ftle_grid_2D

G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_golden.npz"))
res = {}
def rep(k, v):
    res[k] = v
    print(k, v, flush=True)

print("devices", _lib.device_count())
rep("fp64_peak_tflops", _lib.fp64_peak(20000))

x = np.linspace(0, 2, 21); y = np.linspace(0, 1, 11)
f, p, _ = get_predefined_flow("double_gyre")
fo, po, _ = O.get_predefined_flow("double_gyre")
fm = flowmap_grid_2D(f, 0., 8., x, y, p)
rep("golden fm f32 maxdiff", float(np.abs(fm.astype(np.float32) - G["ref_fm"]).max()))
fmo = O.flowmap_grid_2D(fo, 0., 8., x, y, po)
rep("fm vs oracle", float(np.abs(fm - fmo).max()))
fmn, ts = flowmap_n_grid_2D(f, 0., 8., x, y, p, n=4)
rep("golden fm_n f32 maxdiff", float(np.abs(fmn.astype(np.float32) - G["ref_fm_n"]).max()))
fmno, tso = O.flowmap_n_grid_2D(fo, 0., 8., x, y, po, n=4)
rep("fm_n vs oracle", float(np.abs(fmn - fmno).max())); rep("tspan", [ts.tolist(), tso.tolist()])
ft = ftle_grid_2D(G["ref_fm"], 8., x[1], y[1])
rep("golden ftle maxdiff", float(np.abs(ft.astype(np.float32) - G["ref_ftle"]).max()))
rep("ftle seeded vs ref", float(np.abs(ftle_grid_2D(G["ftle_in"], *G["ftle_args"]) - G["ftle_out"]).max()))
rep("ftle seeded masked vs ref", float(np.abs(ftle_grid_2D(G["ftle_in"], *G["ftle_args"], mask=G["ftle_mask"]) - G["ftle_out_masked"]).max()))
vi = get_callable_scalar_linear(((0., 8., 4), (0., 2., 21), (0., 1., 11)), G["ref_vort"])
X, Y = np.meshgrid(x, y, indexing="ij")
lv = lavd_grid_2D(G["ref_fm_n"].astype(np.float64), np.linspace(0, 8, 4), 8., vi, X.ravel(), Y.ravel())
rep("golden lavd maxdiff", float(np.abs(lv.astype(np.float32) - G["ref_lavd"]).max()))

# C1
x = np.linspace(0, 2, 401); y = np.linspace(0, 1, 201)
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
fo, po, _ = O.get_predefined_flow("double_gyre", int_direction=-1.0)
info = {}
t = time.time(); fm = flowmap_grid_2D(f, 0., -10., x, y, p, info=info); t = time.time() - t
fmo, _, sto, stepso, statso = O.flowmap_grid_2D(fo, 0., -10., x, y, po, full=True)
d = np.abs(fm - fmo).max(axis=-1)
same = (info["steps"] == stepso).all(axis=-1)
rep("C1 stats gpu/oracle", [info["stats"].tolist(), statso.tolist()])
rep("C1 step mismatches", int((~same).sum()))
rep("C1 max diff (matching)", float(d[same].max())); rep("C1 max diff (all)", float(d.max()))
rep("C1 wall s", t)

# bickley
fb, pb, dom = get_predefined_flow("bickley_jet")
fbo, pbo, _ = O.get_predefined_flow("bickley_jet")
xb = np.linspace(dom[0][0], dom[0][1], 401); yb = np.linspace(-3, 3, 121)
info = {}
fm = flowmap_grid_2D(fb, 0., 6., xb, yb, pb, info=info)
fmo, _, sto, stepso, statso = O.flowmap_grid_2D(fbo, 0., 6., xb, yb, pbo, full=True)
d = np.abs(fm - fmo).max(axis=-1); same = (info["steps"] == stepso).all(axis=-1)
rep("bickley mismatches", int((~same).sum())); rep("bickley max diff (matching)", float(d[same].max()))
rep("bickley stats", [info["stats"].tolist(), statso.tolist()])

# abc
fa, pa, _ = get_predefined_flow("abc"); fao, pao, _ = O.get_predefined_flow("abc")
rng = np.random.default_rng(1); pts = rng.uniform(0, 2 * np.pi, size=(5000, 3))
info = {}
fm = flowmap(fa, 0., 3., pts, pa, info=info)
fmo, _, sto, stepso, statso = O.flowmap_pts(fao, 0., 3., pts, pao, full=True)
d = np.abs(fm - fmo).max(axis=-1); same = (info["steps"] == stepso).all(axis=-1)
rep("abc mismatches", int((~same).sum())); rep("abc max diff (matching)", float(d[same].max()))

# spline flow
tt = np.linspace(0, 10, 21); xs = np.linspace(0, 2, 41); ys = np.linspace(0, 1, 31)
Tm, Xm, Ym = np.meshgrid(tt, xs, ys, indexing="ij")
a = 0.25 * np.sin(0.2 * np.pi * Tm); b = 1 - 2 * a; ff = a * Xm ** 2 + b * Xm
U = -np.pi * 0.1 * np.sin(np.pi * ff) * np.cos(np.pi * Ym); V = np.pi * 0.1 * np.cos(np.pi * ff) * np.sin(np.pi * Ym) * (2 * a * Xm + b)
grid, Cu, Cv = get_interp_arrays_2D(tt, xs, ys, U, V)
grido, Cuo, Cvo = O.get_interp_arrays_2D(tt, xs, ys, U, V)
rep("prefilter vs oracle", [float(np.abs(Cu - Cuo).max()), float(np.abs(Cv - Cvo).max())])
for mode in ("constant", "linear", "nearest"):
    fs = get_flow_2D(grid, Cu, Cv, extrap_mode=mode); fso = O.get_flow_2D(grido, Cuo, Cvo, extrap_mode=mode)
    xg = np.linspace(0.05, 1.95, 101); yg = np.linspace(0.05, 0.95, 51)
    info = {}
    fm = flowmap_grid_2D(fs, 0., 8., xg, yg, np.array([1.0]), info=info)
    fmo, _, sto, stepso, statso = O.flowmap_grid_2D(fso, 0., 8., xg, yg, np.array([1.0]), full=True)
    d = np.abs(fm - fmo).max(axis=-1); same = (info["steps"] == stepso).all(axis=-1)
    rep("spline %s mismatches/maxdiff/all" % mode, [int((~same).sum()), float(d[same].max()), float(d.max())])

# timing: DG 4096^2 and 8192^2 device resident
import torch
for n in (2048, 4096, 8192):
    x = torch.linspace(0, 2, n, dtype=torch.float64, device="cuda"); y = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
    info = {}
    for it in range(2):
        torch.cuda.synchronize(); t = time.time()
        fmd, ftd = flowmap_ftle_grid_2D(f, 0., -10., x, y, p, 2.0 / (n - 1), 1.0 / (n - 1), device_out=True, info=info if it == 0 else None)
        torch.cuda.synchronize(); t = time.time() - t
    rep("DG %d^2 s, Mpts/s" % n, [t, n * n / t / 1e6]); rep("stats", info["stats"].tolist())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ft2 = ftle_grid_2D(fmd, -10., 2.0 / (n - 1), 1.0 / (n - 1)); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); rep("ftle %d^2 ms, GB/s" % n, [ms, 24.0 * n * n / ms / 1e6])
json.dump(res, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "probe.json"), "w"), indent=1)
