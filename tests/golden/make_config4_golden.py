#!/usr/bin/env python
"""Frozen outputs of the REAL reference for the config-4 call chain
(examples/elliptic_lcs/plot_qge_elliptic_lcs.py:47-88): get_interp_arrays_2D -> get_flow_2D /
get_callable_2D -> flowmap_n_grid_2D -> curl_func_tspan -> get_interp_arrays_scalar ->
get_callable_scalar -> lavd_grid_2D, run from /root/reference/src over oracle/shims
-> tests/golden/config4_golden.npz.

    python tests/golden/make_config4_golden.py

The reference's unmodified Python executes (its numba callables, prange loops, Simpson rule); the
shims supply numbalsoda.dop853 and interpolation.splines (oracle/shims/README.md)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

BODY = r'''
import numpy as np, sys
from numbacs.flows import (get_interp_arrays_2D, get_flow_2D, get_callable_2D, get_interp_arrays_scalar,
                           get_callable_scalar)
from numbacs.integration import flowmap_n_grid_2D
from numbacs.diagnostics import lavd_grid_2D
from numbacs.utils import curl_func_tspan
G = {}
rng = np.random.default_rng(4)
t, x, y = np.linspace(0, 1, 9), np.linspace(0, 1, 25), np.linspace(0, 2, 41)
T_, X, Y = np.meshgrid(t, x, y, indexing="ij")
psi = (0.05 * np.sin(np.pi * X) * np.sin(np.pi * Y / 2) * np.cos(3 * T_ + 0.3)
       + 0.03 * np.sin(2 * np.pi * X) * np.sin(np.pi * Y) * np.cos(5 * T_ + 1.1))
dx, dy = x[1] - x[0], y[1] - y[0]
U, V = -np.gradient(psi, dy, axis=2), np.gradient(psi, dx, axis=1)
G["t"], G["x"], G["y"], G["U"], G["V"] = t, x, y, U, V
grid, Cu, Cv = get_interp_arrays_2D(t, x, y, U, V)
funcptr = get_flow_2D(grid, Cu, Cv, extrap_mode="linear")
vel = get_callable_2D(grid, Cu, Cv, extrap_mode="linear")
pts = np.column_stack((rng.uniform(0, 1, 40), rng.uniform(0.05, 0.95, 40), rng.uniform(0.1, 1.9, 40)))
G["vel_pts"] = pts
G["vel_vals"] = np.array([vel(p) for p in pts])
velt = get_callable_2D(grid, Cu, Cv, extrap_mode="linear", return_type="tuple")
G["vel_tuple0"] = np.array(velt(pts[0]))
# spherical = 1 callable on the same coefficients (the scaling only)
vs = get_callable_2D(grid, Cu, Cv, spherical=1, extrap_mode="linear", r=3.5)
G["vel_sph_vals"] = np.array([vs(p) for p in pts])
xg, yg = np.linspace(0.1, 0.9, 12), np.linspace(0.2, 1.8, 15)
params = np.array([1.0])
n = 9
fmn, tspan = flowmap_n_grid_2D(funcptr, 0.3, 0.4, xg, yg, params, n=n)
G["xg"], G["yg"], G["n"], G["fmn"], G["tspan"] = xg, yg, np.array([n]), fmn, tspan
vort = curl_func_tspan(vel, tspan, xg, yg, h=1e-3)
G["vort"] = vort
gw, Cw = get_interp_arrays_scalar(tspan, xg, yg, vort)
w = get_callable_scalar(gw, Cw)
Xg, Yg = np.meshgrid(xg, yg, indexing="ij")
G["lavd"] = lavd_grid_2D(fmn, tspan, 0.4, w, Xg.ravel(), Yg.ravel())
np.savez_compressed(sys.argv[1], **G)
print("wrote", sys.argv[1], {k: v.shape for k, v in G.items()})
'''


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from run_reference_tests import env_with_shims
    out = os.path.join(HERE, "config4_golden.npz")
    subprocess.check_call([sys.executable, "-c", BODY, out], env=env_with_shims(), cwd="/tmp")


if __name__ == "__main__":
    main()
