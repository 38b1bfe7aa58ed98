#!/usr/bin/env python
"""SHA-1 of flow maps (and per-particle step counts) that every lane-mapping / launch-shape variant
of the kernels must reproduce bit for bit: double gyre 401 x 201, Bickley 333 x 77 and 701 x 203
(the second is large enough for the queue kernels), masked double gyre 61 x 35 and 301 x 251,
a 70 000-point Bickley list.
    [B200CS_LIB=...] python tools/grid_hash.py"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from numbacs_b200 import _lib
from numbacs_b200.flows import get_predefined_flow
from numbacs_b200.integration import flowmap, flowmap_grid_2D


def digest(fm, info):
    h = hashlib.sha1(np.ascontiguousarray(fm).tobytes())
    h.update(np.ascontiguousarray(info["steps"]).tobytes())
    h.update(np.ascontiguousarray(info["status"]).tobytes())
    h.update(np.asarray(info["stats"], dtype=np.int64).tobytes())
    return h.hexdigest()[:10]


out = []
f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
fb, pb, dom = get_predefined_flow("bickley_jet")
rng = np.random.default_rng(1)
info = {}
out.append(digest(flowmap_grid_2D(f, 0.0, -10.0, np.linspace(0, 2, 401), np.linspace(0, 1, 201), p, info=info), info))
for nx, ny in ((333, 77), (701, 203)):
    info = {}
    out.append(digest(flowmap_grid_2D(fb, 0.0, 6.0, np.linspace(dom[0][0], dom[0][1], nx), np.linspace(-3, 3, ny), pb,
                                      info=info), info))
for nx, ny, T in ((61, 35, -4.0), (301, 251, -6.0)):
    mask = rng.random((nx, ny)) < 0.3
    info = {}
    out.append(digest(flowmap_grid_2D(f, 0.0, T, np.linspace(0, 2, nx), np.linspace(0, 1, ny), p, mask=mask,
                                      info=info), info))
pts = np.column_stack((rng.uniform(dom[0][0], dom[0][1], 70000), rng.uniform(-3, 3, 70000)))
info = {}
out.append(digest(flowmap(fb, 0.0, 6.0, pts, pb, info=info), info))
print(f"lib={os.path.basename(_lib.LIB_PATH)} hashes: " + " ".join(out))
