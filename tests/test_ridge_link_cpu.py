"""Host-side ridge linking (csrc/ridge_link.cu: b200cs_link_ridge_pts, b200cs_order_ridges) against
frozen outputs of the REAL reference code (numbacs/extraction/ridges.py:418-603, 720-1054, run
from /root/reference/src by tests/golden/make_ordered_ridges_golden.py) and the reference's own
pickled golden (tests/testing_data/ordered_ridges.pkl, tests/test_extraction.py:23-29).

These entries are host code: they run without a GPU, fed with the stored per-pixel arrays of
_ftle_ridge_pts_connect.  The GPU test (test_gpu_tensor_ridges.py) runs the whole
ftle_ordered_ridges pipeline with the per-pixel stage on the device."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["ref", "dg_a", "dg_b", "dg_c", "rnd_a", "rnd_b", "rnd_c"]


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(ROOT, "tests", "golden", "ordered_ridges_golden.npz"))


def vp(a):
    return C.c_void_p(a.ctypes.data)


def link(lib, G, tag):
    r_pts, r_vec, sdd = (np.ascontiguousarray(G[f"{tag}_{k}"]) for k in ("r_pts", "r_vec", "sdd"))
    nx, ny = G[tag + "_f"].shape
    dist_tol, ang, mrp, thr, pct, c, h = G[tag + "_args"]
    cap = int((sdd < 0).sum())
    ccap = cap // 2 + 1
    linked, rl = np.empty((max(cap, 1), 2)), np.empty((ccap, 2), np.int32)
    ep, tv = np.empty((2 * ccap, 3)), np.empty((2 * ccap, 2))
    counts = np.zeros(2, np.int64)
    rc = lib.b200cs_link_ridge_pts(vp(r_pts), vp(r_vec), vp(sdd), nx, ny, h, c, thr, vp(linked), cap, vp(rl),
                                   vp(ep), vp(tv), ccap, vp(counts))
    assert rc == 0, lib.b200cs_last_error()
    n, k = int(counts[0]), int(counts[1])
    return linked[:n], rl[:k], ep[:2 * k], tv[:2 * k]


@pytest.mark.parametrize("tag", CASES)
def test_linked_ridge_pts_matches_reference(lib, G, tag):
    linked, rl, ep, tv = link(lib, G, tag)
    assert np.array_equal(rl, G[tag + "_ridge_len"])
    assert np.array_equal(linked, G[tag + "_linked"])          # the same points in the same order, bit for bit
    assert np.array_equal(ep, G[tag + "_endpoints"])
    assert np.allclose(tv, G[tag + "_tanvecs"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("tag", CASES)
def test_ordered_ridges_match_reference(lib, G, tag):
    linked, rl, ep, tv = (np.ascontiguousarray(G[f"{tag}_{k}"]) for k in ("linked", "ridge_len", "endpoints", "tanvecs"))
    dist_tol, ang, mrp, thr, pct, c, h = G[tag + "_args"]
    k = len(rl)
    out, offs, n_out = np.empty((max(len(linked), 1), 2)), np.zeros(k + 1, np.int64), np.zeros(1, np.int64)
    rc = lib.b200cs_order_ridges(vp(linked), len(linked), vp(rl), vp(ep), vp(tv), k, dist_tol, ang, int(mrp),
                                 vp(out), vp(offs), vp(n_out))
    assert rc == 0, lib.b200cs_last_error()
    m = int(n_out[0])
    lens = np.diff(offs[:m + 1])
    assert np.array_equal(lens, G[tag + "_ordered_len"])
    assert np.array_equal(out[:offs[m]], G[tag + "_ordered_cat"])
    if tag == "ref":    # the reference's own pickled golden
        assert np.array_equal(lens, G["ref_pkl_ordered_len"])
        assert np.allclose(out[:offs[m]], G["ref_pkl_ordered_cat"])


def test_link_sizes_only_and_errors(lib, G):
    r_pts, r_vec, sdd = (np.ascontiguousarray(G[f"dg_a_{k}"]) for k in ("r_pts", "r_vec", "sdd"))
    nx, ny = G["dg_a_f"].shape
    counts = np.zeros(2, np.int64)
    # no room: sizes come back, nothing is written
    assert lib.b200cs_link_ridge_pts(vp(r_pts), vp(r_vec), vp(sdd), nx, ny, 0.0125, 1.0, 10.0, None, 0, None, None,
                                     None, 0, vp(counts)) == 0
    assert counts[0] == len(G["dg_a_linked"]) and counts[1] == len(G["dg_a_ridge_len"])
    assert lib.b200cs_link_ridge_pts(None, vp(r_vec), vp(sdd), nx, ny, 0.0125, 1.0, 10.0, None, 0, None, None,
                                     None, 0, vp(counts)) != 0
    assert lib.b200cs_link_ridge_pts(vp(r_pts), vp(r_vec), vp(sdd), 3, 3, 0.0125, 1.0, 10.0, None, 0, None, None,
                                     None, 0, vp(counts)) != 0
    # an empty field has no curves and no ridges
    z = np.zeros(nx * ny)
    assert lib.b200cs_link_ridge_pts(vp(-np.ones((nx * ny, 3))), vp(np.zeros((nx * ny, 2))), vp(z), nx, ny, 0.0125,
                                     1.0, 0.0, None, 0, None, None, None, 0, vp(counts)) == 0
    assert counts[0] == 0 and counts[1] == 0
    n_out = np.ones(1, np.int64)
    assert lib.b200cs_order_ridges(None, 0, None, None, None, 0, 0.1, 0.7, 5, None, None, vp(n_out)) == 0 and n_out[0] == 0
