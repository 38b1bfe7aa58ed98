#!/usr/bin/env python
"""Config 2 (bickley_jet 2001 x 601, T = +6): device-timed flow map + FTLE and the test-suite's
parity figures against the oracle on a 4000-particle sample.   [B200CS_LIB=...] python tests/perf/time_bickley.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import oracle as O
from numbacs_b200 import _lib
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import flowmap_grid_2D, flowmap
from numbacs_b200.diagnostics import ftle_grid_2D
fb, pb, dom = get_predefined_flow("bickley_jet")
fbo, pbo, _ = O.get_predefined_flow("bickley_jet")
xb, yb = np.linspace(dom[0][0], dom[0][1], 2001), np.linspace(-3, 3, 601)
xd, yd = torch.tensor(xb, device="cuda"), torch.tensor(yb, device="cuda")
ts = []
for r in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ft = ftle_grid_2D(flowmap_grid_2D(fb, 0.0, 6.0, xd, yd, pb), 6.0, xb[1] - xb[0], yb[1] - yb[0]); e1.record()
    torch.cuda.synchronize()
    if r: ts.append(e0.elapsed_time(e1))
rng = np.random.default_rng(0)
pts = np.column_stack((rng.uniform(dom[0][0], dom[0][1], 4000), rng.uniform(-3, 3, 4000)))
info = {}
g = flowmap(fb, 0.0, 6.0, pts, pb, info=info)
o, _, _, steps_o, _ = O.flowmap_pts(fbo, 0.0, 6.0, pts, pbo, full=True)
d = (np.abs(g - o) / np.array([dom[0][1], 6.0])).max(-1)
same = (info["steps"] == steps_o).all(-1)
print(f"lib={os.path.basename(_lib.LIB_PATH)}: {min(ts):.3f} ms = {2001 * 601 / min(ts) / 1e3:.1f} M pts/s; "
      f"mismatches {int((~same).sum())}/4000, max rel dx (matching) {d[same].max():.2e}, median {np.median(d):.2e}, p99 {np.percentile(d, 99):.2e}")
