#!/usr/bin/env python
"""the reference's own CPU path timed on BASELINE configs 1-4, on this box's
host cores -- the UNMODIFIED reference package (baseline/_ref, or /root/reference/src where it exists)
over oracle/shims, exactly the import bench.py's reference arm uses.

Protocol (examples/time_series/plot_dg_numbacs_vs_scipy.py:170-203): one warm-up call (numba JIT)
excluded, then best of 3 `time.perf_counter()` runs of flowmap_grid_2D + ftle_grid_2D (config 4:
flowmap_n_grid_2D + lavd_grid_2D).  Configs whose full size would take minutes per run are timed on a
contiguous block of rows of the same grid, sized for ~4 s per run, and reported in points/s with the
sample stated (config 4 additionally because the reference materialises the [nx, ny, 601, 2] trajectory
array: 10 GB at full size).  The spline coefficients of configs 3 and 4 are prefiltered on the GPU by
numbacs_b200 (setup, not part of the timed path) and handed to the reference's get_flow_2D as arrays.

    python tests/perf/cpu_configs.py > gpurun_out/cpu_configs.json        (needs a GPU only for that setup)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

import bench

R = bench.reference_over_shims()
if R is None:
    print(json.dumps({"unavailable": "reference package or numba not importable on this box"}))
    sys.exit(0)
import numbacs.diagnostics as rdiag   # noqa: E402  (the reference, on sys.path now)
import numbacs.flows as rflows        # noqa: E402
import numbacs.integration as rint    # noqa: E402

TARGET_S = 4.0
res = {"cores": R["threads"], "host_cpus": os.cpu_count(), "kind": "reference-shim", "source": R["src"],
       "protocol": "warm-up excluded, best of 3, flowmap_grid_2D + ftle_grid_2D (C4: flowmap_n_grid_2D + lavd_grid_2D)"}


def best_of(fn, reps=3):
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


def timed_rows(name, run, nx, ny, full_note):
    """run(i0, rows) integrates rows [i0, i0 + rows) of the config's grid; rows sized for ~TARGET_S."""
    probe = max(4, min(nx, 16))
    i0 = nx // 4
    run(i0, probe)                                  # JIT
    t = best_of(lambda: run(i0, probe), reps=1)
    rows = int(max(probe, min(nx, TARGET_S / t * probe)))
    if nx * ny <= 250_000:
        rows = nx   # small configs run whole
    if rows >= nx:
        rows, i0 = nx, 0
    else:
        i0 = min(i0, nx - rows)
    sec = best_of(lambda: run(i0, rows))
    res[name] = {"points_per_s": rows * ny / sec, "seconds": sec,
                 "sample": (f"the whole {nx} x {ny} grid" if rows == nx else
                            f"{rows} contiguous rows x {ny} columns of the {nx} x {ny} grid ({rows * ny} particles)"),
                 "note": full_note}
    sys.stderr.write(f"{name}: {res[name]['points_per_s'] / 1e6:.3f} M points/s ({res[name]['sample']}, {sec:.2f} s)\n")


# ---- C1: double gyre 401 x 201, T = -10
f, p, _ = rflows.get_predefined_flow("double_gyre", int_direction=-1.0)
x, y = np.linspace(0, 2, 401), np.linspace(0, 1, 201)
timed_rows("C1 double_gyre 401x201 T=-10",
           lambda i0, r: rdiag.ftle_grid_2D(rint.flowmap_grid_2D(f, 0.0, -10.0, np.ascontiguousarray(x[i0:i0 + r]), y, p),
                                            -10.0, x[1] - x[0], y[1] - y[0]), 401, 201, "")

# ---- C2: Bickley jet 2001 x 601, T = +6
fb, pb, dom = rflows.get_predefined_flow("bickley_jet")
xb, yb = np.linspace(dom[0][0], dom[0][1], 2001), np.linspace(-3, 3, 601)
timed_rows("C2 bickley_jet 2001x601 T=+6",
           lambda i0, r: rdiag.ftle_grid_2D(rint.flowmap_grid_2D(fb, 0.0, 6.0, np.ascontiguousarray(xb[i0:i0 + r]), yb, pb),
                                            6.0, xb[1] - xb[0], yb[1] - yb[0]), 2001, 601, "")

# ---- C3 / C4 need spline coefficients: prefiltered on the GPU (setup), evaluated by the reference on the CPU
try:
    import torch
    from numbacs_b200.flows import get_interp_arrays_2D, get_interp_arrays_scalar
    have_gpu = torch.cuda.is_available()
except Exception:   # noqa: BLE001
    have_gpu = False
if have_gpu:
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
    Cu_h, Cv_h = Cu.cpu().numpy(), Cv.cpu().numpy()
    del U, V, Cu, Cv
    torch.cuda.empty_cache()
    fs = rflows.get_flow_2D(grid, Cu_h, Cv_h, spherical=1, extrap_mode="linear")
    lonf, latf = np.arange(-100, 35 + 0.1, 0.2), np.arange(-5, 45 + 0.1, 0.2)
    pm = np.array([-1.0])
    timed_rows("C3 MERRA-shaped spline spherical=1 676x251 T=-72h",
               lambda i0, r: rdiag.ftle_grid_2D(rint.flowmap_grid_2D(fs, 360.0, -72.0, np.ascontiguousarray(lonf[i0:i0 + r]),
                                                                     latf, pm), -72.0, 0.2, 0.2),
               len(lonf), len(latf), "synthetic 720 x 576 x 361 field; coefficients prefiltered on the GPU (setup)")
    del Cu_h, Cv_h

    # ---- C4: LAVD, QGE-shaped field (257 x 513 x 101), 1024 x 1024 particles, n = 601
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
    gw, Cw = get_interp_arrays_scalar(tq, xq, yq, vort)
    host = lambda a: np.ascontiguousarray(a.cpu().numpy() if hasattr(a, "cpu") else a)   # noqa: E731
    fq = rflows.get_flow_2D(gq, host(Cuq), host(Cvq), extrap_mode="linear")
    w = rflows.get_callable_scalar(gw, host(Cw), extrap_mode="linear")
    xp, yp = np.linspace(0.02, 0.98, 1024), np.linspace(0.02, 1.98, 1024)
    one = np.array([1.0])

    def c4(i0, r):
        xs = np.ascontiguousarray(xp[i0:i0 + r])
        fmn, ts = rint.flowmap_n_grid_2D(fq, 0.5, 0.3, xs, yp, one, n=601)
        X, Y = np.meshgrid(xs, yp, indexing="ij")
        return rdiag.lavd_grid_2D(fmn, ts, 0.3, w, X.ravel(), Y.ravel())

    timed_rows("C4 LAVD 1024x1024 n=601 QGE-shaped", c4, 1024, 1024,
               "flowmap_n_grid_2D + lavd_grid_2D; the spatial-mean vorticity is taken over the sampled rows "
               "(the reference's own semantics for the points it is given)")
else:
    res["C3/C4"] = "skipped: no GPU on this box for the spline-coefficient setup"

print(json.dumps(res, indent=1))
