#!/usr/bin/env python
"""Parity at BASELINE sizes, product build and parity-calibration (strict) build side by side.

    python tests/perf/parity_floor.py [c1 c2 c3 c5 small] > gpurun_out/parity_floor.json

For each configuration the same particles are integrated by (a) the CPU oracle, (b) the product
library libb200cs.so, (c) libb200cs_strict.so (B200CS_STRICT: reference evaluation order,
separately rounded operations, CUDA libm -- csrc/dop853.cuh).  Reported per pair: step-count
mismatches and max|dx|/L over step-matching particles.  strict-vs-oracle is the distance that is
left when ONLY the libm differs: the floor under any GPU implementation of this path.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import oracle as O
import numbacs_b200 as nb
from numbacs_b200 import _lib
from parity_common import (compare_flowmaps, ftle_rel_l2, bickley_grid, merra_axes, merra_field,
                           merra_particles, c5_sample_rows)

want = set(a for a in sys.argv[1:] if "=" not in a) or {"c1", "c2", "c3", "c5", "small"}
# extra A/B libraries: NAME=path arguments (tools/build_tu_variant.sh), each compared with the oracle too
extra_libs = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
res = {}
_extra_runs = {}


def both(fn):
    """fn() under the product library and under the strict one (and under every extra library)."""
    a = fn()
    with _lib.use_library(_lib.STRICT_LIB_PATH):
        b = fn()
    _extra_runs.clear()
    for nm, path in extra_libs.items():
        with _lib.use_library(os.path.abspath(path)):
            _extra_runs[nm] = fn()
    return a, b


def triple(name, L, gpu, strict, ora, extra=None):
    (fm, st), (fs, ss), (fo, so) = gpu, strict, ora
    r_fast, same_f = compare_flowmaps(fm, st, fo, so, L)
    r_strict, same_s = compare_flowmaps(fs, ss, fo, so, L)
    r_fs, _ = compare_flowmaps(fm, st, fs, ss, L)
    res[name] = {"product_vs_oracle": r_fast, "strict_vs_oracle": r_strict, "product_vs_strict": r_fs}
    if extra:
        res[name].update(extra)
    sys.stderr.write(f"{name}: product {r_fast['step_mismatches']} mism, max {r_fast['max_rel_dx_matching']:.2e} "
                     f"p99 {r_fast['p99_rel_dx']:.2e} | strict {r_strict['step_mismatches']} mism, "
                     f"max {r_strict['max_rel_dx_matching']:.2e} p99 {r_strict['p99_rel_dx']:.2e}\n")
    for nm, (fe, se) in _extra_runs.items():
        r, _ = compare_flowmaps(fe, se, fo, so, L)
        res[name][nm + "_vs_oracle"] = r
        sys.stderr.write(f"    {nm}: {r['step_mismatches']} mism, max {r['max_rel_dx_matching']:.2e} "
                         f"p99 {r['p99_rel_dx']:.2e} over1e-8 {r['over_1e-8_matching']}\n")
    return same_f, same_s


def grid_run(flow_name, direction, t0, T, x, y, rows=None, **flow_kw):
    def run():
        f, p, _ = nb.flows.get_predefined_flow(flow_name, int_direction=direction, **flow_kw)
        info = {}
        fm = nb.integration.flowmap_grid_2D(f, t0, T, x, y, p, info=info)
        st = np.asarray(info["steps"])
        assert (np.asarray(info["status"]) == 1).all()
        return (fm, st) if rows is None else (fm[rows], st[rows])
    return run


if "c1" in want:
    x, y = np.linspace(0, 2, 401), np.linspace(0, 1, 201)
    g, s = both(grid_run("double_gyre", -1.0, 0.0, -10.0, x, y))
    fo, po, _ = O.get_predefined_flow("double_gyre", int_direction=-1.0)
    fmo, _, _, so, _ = O.flowmap_grid_2D(fo, 0.0, -10.0, x, y, po, full=True)
    same_f, same_s = triple("C1 double_gyre 401x201 T=-10", (2.0, 1.0), g, s, (fmo, so))
    dx, dy = x[1] - x[0], y[1] - y[0]
    fto = O.ftle_grid_2D(fmo, -10.0, dx, dy)
    res["C1 double_gyre 401x201 T=-10"]["ftle_rel_l2_product"] = ftle_rel_l2(nb.diagnostics.ftle_grid_2D(g[0], -10.0, dx, dy), fto, same_f)
    res["C1 double_gyre 401x201 T=-10"]["ftle_rel_l2_strict"] = ftle_rel_l2(nb.diagnostics.ftle_grid_2D(s[0], -10.0, dx, dy), fto, same_s)

if "c2" in want:
    x, y = bickley_grid()
    rows = np.array(sorted(set(range(0, 2001, 25)) | set(range(1000, 1016))))
    g, s = both(grid_run("bickley_jet", 1.0, 0.0, 6.0, x, y, rows=rows))
    fo, po, _ = O.get_predefined_flow("bickley_jet")
    fmo, _, _, so, _ = O.flowmap_grid_2D(fo, 0.0, 6.0, x[rows], y, po, full=True)
    triple("C2 bickley_jet 2001x601 T=+6 (97 rows sampled)", (x[-1] - x[0], 6.0), g, s, (fmo, so))
    # a short horizon for scale: the same grid, T = 1
    g, s = both(grid_run("bickley_jet", 1.0, 0.0, 1.0, x, y, rows=rows))
    fmo, _, _, so, _ = O.flowmap_grid_2D(fo, 0.0, 1.0, x[rows], y, po, full=True)
    triple("C2 bickley_jet 2001x601 T=+1 (97 rows sampled)", (x[-1] - x[0], 6.0), g, s, (fmo, so))

if "c3" in want:
    t, lon, lat = merra_axes()
    td, lond, latd = (torch.tensor(v, device="cuda") for v in (t, lon, lat))
    U, V = merra_field(torch, td, lond, latd)
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(t, lon, lat, U, V)
    del U, V
    Cu_h, Cv_h = Cu.cpu().numpy(), Cv.cpu().numpy()
    lonf, latf = merra_particles()
    pm = np.array([-1.0])

    def run():
        fs = nb.flows.get_flow_2D(grid, Cu, Cv, spherical=1, extrap_mode="linear")
        info = {}
        fm = nb.integration.flowmap_grid_2D(fs, 360.0, -72.0, lonf, latf, pm, info=info)
        assert (np.asarray(info["status"]) == 1).all()
        return fm, np.asarray(info["steps"])
    g, s = both(run)
    fso = O.get_flow_2D(grid, Cu_h, Cv_h, spherical=1, extrap_mode="linear")
    t0 = time.time()
    fmo, _, _, so, _ = O.flowmap_grid_2D(fso, 360.0, -72.0, lonf, latf, pm, full=True)
    same_f, same_s = triple("C3 MERRA-shaped spline spherical=1 676x251 T=-72h", (360.0, 180.0), g, s, (fmo, so),
                            {"oracle_s": time.time() - t0})
    fto = O.ftle_grid_2D(fmo, -72.0, 0.2, 0.2)
    res["C3 MERRA-shaped spline spherical=1 676x251 T=-72h"]["ftle_rel_l2_product"] = ftle_rel_l2(
        nb.diagnostics.ftle_grid_2D(g[0], -72.0, 0.2, 0.2), fto, same_f)
    del Cu, Cv, Cu_h, Cv_h
    torch.cuda.empty_cache()

if "c5" in want:
    n = 16384
    x, y = np.linspace(0, 2, n), np.linspace(0, 1, n)
    starts, rpb = c5_sample_rows(n, blocks=8, rows_per_block=6)
    rows = np.concatenate([np.arange(a, a + rpb) for a in starts])
    g, s = both(grid_run("double_gyre", -1.0, 0.0, -10.0, x[rows], y))
    fo, po, _ = O.get_predefined_flow("double_gyre", int_direction=-1.0)
    t0 = time.time()
    fmo, _, _, so, _ = O.flowmap_grid_2D(fo, 0.0, -10.0, x[rows], y, po, full=True)
    triple(f"C5 double_gyre 16384x16384 T=-10 ({len(rows)} rows sampled)", (2.0, 1.0), g, s, (fmo, so),
           {"oracle_s": time.time() - t0, "rows": [int(v) for v in rows]})

if "small" in want:
    # the reduced cases of tests/test_gpu_parity.py whose gates are calibrated on the strict build
    x, y = np.linspace(0, 6.371 * np.pi, 401), np.linspace(-3, 3, 121)
    g, s = both(grid_run("bickley_jet", 1.0, 0.0, 6.0, x, y))
    fo, po, _ = O.get_predefined_flow("bickley_jet")
    fmo, _, _, so, _ = O.flowmap_grid_2D(fo, 0.0, 6.0, x, y, po, full=True)
    triple("bickley_jet 401x121 T=+6", (x[-1], 6.0), g, s, (fmo, so))
    pts = np.random.default_rng(1).uniform(0, 2 * np.pi, size=(3000, 3))

    def run_abc():
        f, p, _ = nb.flows.get_predefined_flow("abc")
        info = {}
        fm = nb.integration.flowmap(f, 0.0, 2.0, pts, p, info=info)
        return fm, np.asarray(info["steps"])
    g, s = both(run_abc)
    fo, po, _ = O.get_predefined_flow("abc")
    fmo, _, _, so, _ = O.flowmap_pts(fo, 0.0, 2.0, pts, po, full=True)
    triple("abc 3000 points T=2", 2 * np.pi, g, s, (fmo, so))

res["_libraries"] = {"product": _lib.library_info()["path"], "strict": _lib.STRICT_LIB_PATH}
print(json.dumps(res, indent=1))
