#!/usr/bin/env python
"""Frozen outputs of the REAL reference linking code (numbacs/extraction/ridges.py:
_ftle_ridge_pts_connect, _linked_ridge_pts, ftle_ordered_ridges), run in this container from
/root/reference/src -> tests/golden/ordered_ridges_golden.npz.

    python tests/golden/make_ordered_ridges_golden.py

ridges.py is loaded without the package __init__ (which needs the `interpolation` package that
is not installed); the functions used here only need numba / numpy.  Cases:
  ref   the reference's own fixture (tests/testing_data/ftle.npy + Cevecs.npy, dist_tol 0.1) and
        its pickled goldens ridges.pkl / ordered_ridges.pkl (tests/test_extraction.py:15-29)
  dg    a double-gyre FTLE field, 161 x 81, T = -10 (flow map from the CPU oracle, Cauchy-Green
        eigen-pairs from the real numbacs.diagnostics.C_eig_2D), with the parameters of
        examples/ftle/plot_dg_ftle_ridges.py:64-72 (sdd_thresh 10, dist_tol 5e-2) and two more sets
  rnd   the seeded field of reference_golden.npz (ridge_f / ridge_ev: noisy, random directions:
        many short curves, every branch of the end-point matcher gets exercised)
"""
import importlib.util
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, ROOT)


def load_ridges():
    import numbacs  # noqa: F401  (package __init__ is light; extraction's is not)
    pkg = types.ModuleType("numbacs.extraction")
    pkg.__path__ = [os.path.join(REF, "src", "numbacs", "extraction")]
    sys.modules["numbacs.extraction"] = pkg
    spec = importlib.util.spec_from_file_location(
        "numbacs.extraction.ridges", os.path.join(pkg.__path__[0], "ridges.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def pack(G, tag, ridges, f, ev, x, y, dist_tol, ep_tan_ang, min_ridge_pts, sdd_thresh, percentile, c):
    rp, rv, sdd, h = ridges._ftle_ridge_pts_connect(f, ev, x, y, sdd_thresh, percentile)
    lk, rl, ep, tv = ridges._linked_ridge_pts(f, ev, x, y, sdd_thresh, percentile, c)
    od = ridges.ftle_ordered_ridges(f, ev, x, y, dist_tol, ep_tan_ang, min_ridge_pts, sdd_thresh, percentile, c)
    G[tag + "_f"], G[tag + "_ev"], G[tag + "_x"], G[tag + "_y"] = f, ev, x, y
    G[tag + "_args"] = np.array([dist_tol, ep_tan_ang, min_ridge_pts, sdd_thresh, percentile, c, h])
    G[tag + "_r_pts"], G[tag + "_r_vec"], G[tag + "_sdd"] = rp, rv, sdd
    G[tag + "_linked"], G[tag + "_ridge_len"], G[tag + "_endpoints"], G[tag + "_tanvecs"] = lk, rl, ep, tv
    G[tag + "_ordered_cat"] = np.concatenate(od) if od else np.zeros((0, 2))
    G[tag + "_ordered_len"] = np.array([len(r) for r in od], np.int64)
    print(tag, "ridge pixels", int((sdd < 0).sum()), "curves", len(rl), "ordered ridges", len(od),
          "lengths", [len(r) for r in od][:12])


def main():
    ridges = load_ridges()
    from numbacs.diagnostics import C_eig_2D, ftle_from_eig
    import oracle as O
    G = {}
    td = os.path.join(REF, "tests", "testing_data")
    x, y = np.linspace(0, 2, 21), np.linspace(0, 1, 11)
    f = np.load(os.path.join(td, "ftle.npy")).astype(np.float64)
    ev = np.ascontiguousarray(np.load(os.path.join(td, "Cevecs.npy")).astype(np.float64)[:, :, :, 1])
    pack(G, "ref", ridges, f, ev, x, y, 1e-1, np.pi / 4, 5, 0.0, 0, 1.0)
    with open(os.path.join(td, "ordered_ridges.pkl"), "rb") as fh:
        od = pickle.load(fh)
    G["ref_pkl_ordered_cat"], G["ref_pkl_ordered_len"] = np.concatenate(od), np.array([len(r) for r in od], np.int64)

    x, y = np.linspace(0, 2, 161), np.linspace(0, 1, 81)
    fl, p, _ = O.get_predefined_flow("double_gyre", int_direction=-1.0)
    fm = O.flowmap_grid_2D(fl, 0.0, -10.0, x, y, p)
    vals, vecs = C_eig_2D(fm, x[1] - x[0], y[1] - y[0])
    f = ftle_from_eig(vals[:, :, 1], -10.0)
    ev = np.ascontiguousarray(vecs[:, :, :, 1])
    pack(G, "dg_a", ridges, f, ev, x, y, 5e-2, np.pi / 4, 5, 10.0, 0, 1.0)
    pack(G, "dg_b", ridges, f, ev, x, y, 1e-1, np.pi / 3, 8, 0.0, 50, 0.5)
    pack(G, "dg_c", ridges, f, ev, x, y, 2e-2, np.pi / 6, 2, 30.0, 0, 2.0)

    R = np.load(os.path.join(HERE, "reference_golden.npz"))
    f, ev, x, y = R["ridge_f"], R["ridge_ev"], R["ridge_x"], R["ridge_y"]
    pack(G, "rnd_a", ridges, f, ev, x, y, 1e-1, np.pi / 4, 5, 0.0, 0, 1.0)
    pack(G, "rnd_b", ridges, f, ev, x, y, 2e-1, np.pi / 2, 3, 5.0, 0, 1.0)
    pack(G, "rnd_c", ridges, f, ev, x, y, 6e-2, np.pi / 3, 1, 0.0, 60, 1.0)
    out = os.path.join(HERE, "ordered_ridges_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
