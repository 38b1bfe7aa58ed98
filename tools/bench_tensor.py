#!/usr/bin/env python
"""Device-timed throughput of the flow-map consumers next to ftle_grid_2D:
C_eig_2D, ftle_from_eig, ftle_ridge_pts (config 5's ridge tail), flowmap_aux_grid_2D + C_eig_aux_2D.

    python tools/bench_tensor.py [n=8192] [reps=5]  ->  one JSON line

Everything stays on the device (torch CUDA tensors in, CUDA tensors out); times are CUDA events on
the launching stream, best of `reps` after one warm-up; GB/s figures use ALGORITHMIC bytes per
pixel (stated per entry), peak = MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from numbacs_b200.diagnostics import C_eig_2D, C_eig_aux_2D, ftle_from_eig, ftle_grid_2D
from numbacs_b200.extraction import ftle_ridge_pts
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import flowmap_aux_grid_2D, flowmap_grid_2D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6531.9))


def timed(fn):
    fn()
    best = float("inf")
    out = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


x = torch.linspace(0, 2, n, dtype=torch.float64, device="cuda")
y = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
dx, dy = 2.0 / (n - 1), 1.0 / (n - 1)
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
res = {"n": n, "pixels": n * n, "hbm_peak_gbs": hbm}

t_fm, fm = timed(lambda: flowmap_grid_2D(f, 0.0, -10.0, x, y, p, device_out=True))
res["flowmap_ms"] = t_fm
t, (vals, vecs) = timed(lambda: C_eig_2D(fm, dx, dy))
res["C_eig_2D"] = {"ms": t, "bytes_per_pixel": 64, "gbs": 64 * n * n / t / 1e6, "frac_hbm": 64 * n * n / t / 1e6 / hbm}
t, ftle = timed(lambda: ftle_from_eig(vals[:, :, 1], -10.0))
res["ftle_from_eig"] = {"ms": t, "bytes_per_pixel": 16, "gbs": 16 * n * n / t / 1e6}
t, ftle2 = timed(lambda: ftle_grid_2D(fm, -10.0, dx, dy))
res["ftle_grid_2D"] = {"ms": t, "bytes_per_pixel": 24, "gbs": 24 * n * n / t / 1e6, "frac_hbm": 24 * n * n / t / 1e6 / hbm}
res["ftle_two_routes_rel_l2"] = float(torch.linalg.norm(ftle - ftle2) / torch.linalg.norm(ftle2))
t, r = timed(lambda: ftle_ridge_pts(ftle, vecs[:, :, :, 1], x, y, sdd_thresh=10.0, percentile=0))
# two calls (count, then fill) x (detect + compact): f (8 B) + eigvec (16 B) read per pass
res["ftle_ridge_pts"] = {"ms": t, "n_ridge_pts": int(r.shape[0]), "passes": 3, "bytes_per_pixel": 3 * 24,
                         "gbs": 72 * n * n / t / 1e6}
t, r90 = timed(lambda: ftle_ridge_pts(ftle, vecs[:, :, :, 1], x, y, sdd_thresh=10.0, percentile=90))
res["ftle_ridge_pts_p90"] = {"ms": t, "n_ridge_pts": int(r90.shape[0])}
res["ridge_tail_ms"] = res["C_eig_2D"]["ms"] + res["ftle_from_eig"]["ms"] + res["ftle_ridge_pts"]["ms"]
res["ridge_tail_share_of_flowmap"] = res["ridge_tail_ms"] / t_fm
del vals, vecs, ftle, ftle2, r, r90
torch.cuda.empty_cache()

na = min(n, 4096)   # 5 particles per cell
xa, ya = x[:na].contiguous(), y[:na].contiguous()
t, fa = timed(lambda: flowmap_aux_grid_2D(f, 0.0, -10.0, xa, ya, p, device_out=True))
res["flowmap_aux_grid_2D"] = {"n": na, "ms": t, "particles": 5 * na * na - 8 * (2 * na - 2) // 2,
                              "Mparticles_per_s": 5 * na * na / t / 1e3}
t, _ = timed(lambda: C_eig_aux_2D(fa, dx, dy))
res["C_eig_aux_2D"] = {"n": na, "ms": t, "bytes_per_pixel": 80 + 48, "gbs": 128 * na * na / t / 1e6}
print(json.dumps(res))
