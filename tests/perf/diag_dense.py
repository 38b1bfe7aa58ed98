#!/usr/bin/env python
"""Diagnostic: which particles of the n = 50 dense-output case (tests/test_gpu_parity.py::
test_flowmap_n_dense_output) take a different step sequence than the CPU oracle, and by how much
the rows differ.   [B200CS_LIB=...] python tests/perf/diag_dense.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import oracle
import numbacs_b200 as nb
from numbacs_b200 import _lib

x, y = np.linspace(0, 2, 64), np.linspace(0, 1, 33)
f, p, _ = nb.flows.get_predefined_flow("double_gyre")
fo, po, _ = oracle.get_predefined_flow("double_gyre")
info = {}
fmn, ts = nb.integration.flowmap_n_grid_2D(f, 1.0, 9.0, x, y, p, info=info)
fmno, tso, _, steps_o, stats_o = oracle.flowmap_n_grid_2D(fo, 1.0, 9.0, x, y, po, full=True)
same = (info["steps"] == steps_o).all(axis=-1)
d = np.abs(fmn - fmno).max(axis=(-1, -2))
dc = np.abs(fmn - fmno).max(axis=(0, 1, 2))
print(f"   per-component max |d|: x {dc[0]:.3e} (L = 2), y {dc[1]:.3e} (L = 1); "
      f"count > 5e-9: {int((np.abs(fmn - fmno).max(axis=(-1, -2)) > 5e-9).sum())} of {same.size}")
print(f"lib={os.path.basename(_lib.LIB_PATH)}: mismatches {int((~same).sum())}, max|d| matching {d[same].max():.2e}, "
      f"max|d| all {d.max():.2e}")
for (i, j) in np.argwhere(~same):
    print(f"   ({i},{j}) x={x[i]:.4f} y={y[j]:.4f} gpu {info['steps'][i, j]} oracle {steps_o[i, j]} |d|={d[i, j]:.2e}")
# the final-time kernel on the same case
info2 = {}
fm = nb.integration.flowmap_grid_2D(f, 1.0, 9.0, x, y, p, info=info2)
fmo, _, _, s2, _ = oracle.flowmap_grid_2D(fo, 1.0, 9.0, x, y, po, full=True)
same2 = (info2["steps"] == s2).all(axis=-1)
print(f"   final-time kernel: mismatches {int((~same2).sum())} at {np.argwhere(~same2).tolist()}, "
      f"max|d| {np.abs(fm - fmo).max():.2e}")
