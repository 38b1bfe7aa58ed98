#!/usr/bin/env python
"""The reference's FTLE time-series workload (examples/time_series/plot_dg_time_series.py: double
gyre 201 x 101, T = 16, n = 200 frames one time unit apart), device-resident, CUDA-event timed:
  loop     : flowmap_grid_2D + ftle_grid_2D per frame (what the example's "standard" branch does)
  batched  : flowmap_grid_2D_series + ftle_grid_2D_series, all frames in one launch each
  composed : flowmap_composition_initial + n-1 flowmap_composition_step (+ ftle per frame)
  composed_batched : flowmap_composition_series + ftle_grid_2D_series (three launches in total)
    python tools/time_series.py [nx=201] [ny=101] [n=200]  -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import (flowmap_grid_2D, flowmap_grid_2D_series, flowmap_composition_initial,
                                      flowmap_composition_step, flowmap_composition_series)
from numbacs_b200.diagnostics import ftle_grid_2D, ftle_grid_2D_series

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 201
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 101
n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
f, p, dom = get_predefined_flow("double_gyre", int_direction=1.0)
x = torch.linspace(dom[0][0], dom[0][1], nx, dtype=torch.float64, device="cuda")
y = torch.linspace(dom[1][0], dom[1][1], ny, dtype=torch.float64, device="cuda")
dx, dy = float(x[1] - x[0]), float(y[1] - y[0])
grid = ((float(x[0]), float(x[-1]), nx), (float(y[0]), float(y[-1]), ny))
t0, T, h = 0.0, 16.0, 1.0
tspan = np.arange(t0, t0 + n * h, h)


def timed(fn, reps=3):
    fn()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def loop():
    return torch.stack([ftle_grid_2D(flowmap_grid_2D(f, float(t), T, x, y, p), T, dx, dy) for t in tspan])


def batched():
    return ftle_grid_2D_series(flowmap_grid_2D_series(f, tspan, T, x, y, p), T, dx, dy)


def composed():
    out = []
    fm0, fms, nT = flowmap_composition_initial(f, t0, T, h, x, y, grid, p)
    out.append(ftle_grid_2D(fm0, T, dx, dy))
    for k in range(1, n):
        fmk, fms = flowmap_composition_step(fms, f, t0 + T + (k - 1) * h, h, nT, x, y, grid, p)
        out.append(ftle_grid_2D(fmk, T, dx, dy))
    return torch.stack(out)


def composed_batched():
    return ftle_grid_2D_series(flowmap_composition_series(f, t0, T, h, n, x, y, grid, p), T, dx, dy)


t_loop, a = timed(loop)
t_bat, b = timed(batched)
t_comp, c = timed(composed)
t_cb, d = timed(composed_batched)
pts = n * nx * ny
print(json.dumps({"grid": [nx, ny], "frames": n, "T": T,
                  "loop_ms": t_loop, "batched_ms": t_bat, "composed_ms": t_comp,
                  "loop_Mpts_per_s": pts / t_loop / 1e3, "batched_Mpts_per_s": pts / t_bat / 1e3,
                  "composed_batched_ms": t_cb, "composed_batched_equals_composed": bool(torch.equal(c, d)),
                  "batched_equals_loop": bool(torch.equal(a, b)),
                  "composed_vs_direct_median_abs_ftle_diff": float((c - a).abs().median())}))
