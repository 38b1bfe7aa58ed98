"""FTLE ridge points -- drop-in for the data-parallel part of ``numbacs.extraction.ridges``.

ftle_ridge_pts (extraction/ridges.py:9-76) and _ftle_ridge_pts_connect (232-318): the per-pixel
sub-pixel ridge test (f > f_min, second directional derivative of f along the dominant
Cauchy-Green eigenvector below -sdd_thresh, first-order root inside the pixel) as a CUDA stencil
pass followed by an order-preserving stream compaction, so the points come back in the raveled
pixel order of the reference's ``r_pts[ridge_bool, :]``.  ``np.percentile(f, percentile)`` becomes
two order statistics found by a radix select on the device, combined with numba's interpolation
formula (numba/np/arraymath.py: rank = 1 + (n-1) p/100, lower (1-m) + upper m).

ftle_ridges (ridges.py:175-229) groups the ridge points into 8-connected ridges: the reference
labels ridge_bool with scipy.ndimage.label and gathers each label with a full-grid comparison;
here a union-find kernel labels the ridge pixels on the device and only the (few) ridge points
are grouped on the host.

The serial greedy linking of ridge points into ORDERED curves (_linked_ridge_pts ridges.py:418-603,
ftle_ordered_ridges 720-1054) is inherently sequential; it runs on the host in native code
(csrc/ridge_link.cu, b200cs_link_ridge_pts / b200cs_order_ridges) on the per-pixel arrays the
device kernel behind _ftle_ridge_pts_connect produced.
"""
import ctypes as C
from math import floor, pi

import numpy as np

from . import _lib

__all__ = ["ftle_ridge_pts", "ftle_ridges", "ftle_ordered_ridges", "percentile_value"]


def percentile_value(f, percentile):
    """np.percentile(f, percentile) as numba computes it, with the selection done on the GPU."""
    fa = _lib.arg_in(f)
    n = int(np.prod(fa.obj.shape))
    if n == 0:
        return float("nan")
    p = float(percentile)
    rank = 1 + (n - 1) * (p / 100.0)
    fl = floor(rank)
    m = rank - fl
    k = min(max(int(fl) - 1, 0), n - 1)
    out = np.empty(2, np.float64)
    _lib.check(_lib.load().b200cs_order_stats(fa.ptr, n, k, C.c_void_p(out.ctypes.data),
                                              _lib.current_stream(fa.on_device)))
    if fa.on_device:
        import torch
        torch.cuda.current_stream().synchronize()
    if p == 100:
        return float(out[0])  # k = n - 1: the maximum
    return float(out[0] * (1 - m) + out[1] * m)


def _eigvec_in(eigvec_max, nx, ny):
    """eigvec_max (nx, ny, 2), typically the view eigvecs[:, :, :, 1] of C_eig_2D's output: passed
    in place with (pixel, component) strides when its layout allows it."""
    is_t = _lib._is_torch(eigvec_max)
    a = eigvec_max if is_t else np.asarray(eigvec_max)
    if tuple(int(v) for v in a.shape) != (nx, ny, 2):
        raise ValueError(f"eigvec_max must have shape {(nx, ny, 2)}")
    f64 = (str(a.dtype) == "torch.float64") if is_t else (a.dtype == np.float64)
    if f64:
        si, sj, sc = _lib._elem_strides(a)
        if sj >= 1 and sc >= 1 and si == sj * ny:
            ptr = a.data_ptr() if is_t else a.ctypes.data
            return _lib.Arg(a, C.c_void_p(ptr), bool(is_t and a.is_cuda)), sj, sc
    return _lib.arg_in(a), 2, 1


def _spacing(xa, ya, spacing):
    """dx = x[1] - x[0], dy = y[1] - y[0] (ridges.py:39-40) unless given (row slabs of a larger grid)."""
    if spacing is not None:
        return float(spacing[0]), float(spacing[1])
    return float(xa.obj[1] - xa.obj[0]), float(ya.obj[1] - ya.obj[0])


def _ridge_call(f, eigvec_max, x, y, sdd_thresh, percentile, full, device_out, spacing=None):
    fa, xa, ya = _lib.arg_in(f), _lib.arg_in(x), _lib.arg_in(y)
    if fa.obj.ndim != 2:
        raise ValueError("f must have shape (nx, ny)")
    nx, ny = int(fa.obj.shape[0]), int(fa.obj.shape[1])
    if nx < 2 or ny < 2:
        raise ValueError("the grid needs at least 2 points per axis")
    ev, ps, cs = _eigvec_in(eigvec_max, nx, ny)
    f_min = 0.0 if percentile == 0 else percentile_value(fa.obj, percentile)
    dx, dy = _spacing(xa, ya, spacing)
    dev = bool(device_out or fa.on_device)
    L = _lib.load()
    stream = _lib.current_stream(dev)
    count = np.zeros(1, np.int64)
    cptr = C.c_void_p(count.ctypes.data)
    if full:
        r_pts = _lib.alloc_out((nx * ny, 3), np.float64, dev)
        r_vec = _lib.alloc_out((nx * ny, 2), np.float64, dev)
        sdd = _lib.alloc_out((nx * ny,), np.float64, dev)
        _lib.check(L.b200cs_ftle_ridge_pts(fa.ptr, ev.ptr, ps, cs, nx, ny, xa.ptr, ya.ptr, dx, dy,
                                           float(sdd_thresh), f_min, r_pts.ptr, r_vec.ptr, sdd.ptr,
                                           None, 0, None, stream))
        return r_pts.obj, r_vec.obj, sdd.obj
    # one call = detect + scan + compact into a buffer sized by a guess (ridges are curves: their
    # points are a small fraction of the pixels); only if the guess was too small, a second call
    cap = max(4096, (nx * ny) // 32)
    for _ in range(2):
        pts = _lib.alloc_out((cap, 2), np.float64, dev)
        _lib.check(L.b200cs_ftle_ridge_pts(fa.ptr, ev.ptr, ps, cs, nx, ny, xa.ptr, ya.ptr, dx, dy,
                                           float(sdd_thresh), f_min, None, None, None, pts.ptr, cap,
                                           cptr, stream))
        n = int(count[0])
        if n <= cap:
            break
        cap = n
    return pts.obj[:n].clone() if dev else pts.obj[:n].copy()


def ftle_ridge_pts(f, eigvec_max, x, y, sdd_thresh=0.0, percentile=0, *, device_out=False, spacing=None):
    """Sub-pixel FTLE ridge points -> (k, 2), in raveled pixel order.  `spacing` = (dx, dy)
    overrides x[1] - x[0], y[1] - y[0] (used for row slabs of a larger grid)."""
    return _ridge_call(f, eigvec_max, x, y, sdd_thresh, percentile, False, device_out, spacing)


def _ftle_ridge_pts_connect(f, eigvec_max, x, y, sdd_thresh=0.0, percentile=0, *, device_out=False):
    """Per-pixel ridge data for the linking stage -> (r_pts (nx*ny, 3), r_vec (nx*ny, 2),
    sdd (nx*ny,), h = min(dx, dy))."""
    r_pts, r_vec, sdd = _ridge_call(f, eigvec_max, x, y, sdd_thresh, percentile, True, device_out)
    xs, ys = _lib.arg_in(x).obj, _lib.arg_in(y).obj
    h = min(float(xs[1] - xs[0]), float(ys[1] - ys[0]))
    return r_pts, r_vec, sdd, h


def ftle_ridges(f, eigvec_max, x, y, sdd_thresh=0.0, percentile=0, min_ridge_pts=3):
    """Connected FTLE ridges -> list of (k_i, 2) arrays, one per 8-connected group of ridge pixels
    with at least min_ridge_pts points, in scipy.ndimage.label's (raster) order; points inside a
    ridge are in raveled pixel order, as in the reference."""
    fa, xa, ya = _lib.arg_in(f), _lib.arg_in(x), _lib.arg_in(y)
    if fa.obj.ndim != 2:
        raise ValueError("f must have shape (nx, ny)")
    nx, ny = int(fa.obj.shape[0]), int(fa.obj.shape[1])
    ev, ps, cs = _eigvec_in(eigvec_max, nx, ny)
    f_min = 0.0 if percentile == 0 else percentile_value(fa.obj, percentile)
    dx, dy = _spacing(xa, ya, None)
    L = _lib.load()
    stream = _lib.current_stream(fa.on_device)
    count = np.zeros(1, np.int64)
    cap = max(4096, (nx * ny) // 32)
    for _ in range(2):
        pts = np.empty((cap, 2), np.float64)
        roots = np.empty(cap, np.int64)
        _lib.check(L.b200cs_ftle_ridges(fa.ptr, ev.ptr, ps, cs, nx, ny, xa.ptr, ya.ptr, dx, dy,
                                        float(sdd_thresh), f_min, C.c_void_p(pts.ctypes.data),
                                        C.c_void_p(roots.ctypes.data), cap,
                                        C.c_void_p(count.ctypes.data), stream))
        n = int(count[0])
        if n <= cap:
            break
        cap = n
    pts, roots = pts[:n], roots[:n]
    order = np.argsort(roots, kind="stable")          # ascending root = label order; stable keeps
    sroots = roots[order]                             # the raveled order inside each ridge
    cuts = np.flatnonzero(np.diff(sroots)) + 1
    groups = np.split(order, cuts) if n else []
    return [pts[g] for g in groups if len(g) >= min_ridge_pts]


def _host(a):
    return np.ascontiguousarray(a.cpu().numpy() if _lib._is_torch(a) else a, dtype=np.float64)


def _linked_ridge_pts(f, eigvec_max, x, y, sdd_thresh=0.0, percentile=0, c=1.0):
    """Ridge points ordered and grouped into curves (extraction/ridges.py:418-603) ->
    (linked_ridges_arr (n, 2), ridge_len (k, 2) int32, endpoints (2k, 3), ep_tanvecs (2k, 2)).

    The per-pixel ridge test runs on the GPU (_ftle_ridge_pts_connect); the greedy walk over its
    output is serial and runs on the host (b200cs_link_ridge_pts)."""
    r_pts, r_vec, sdd, h = _ftle_ridge_pts_connect(f, eigvec_max, x, y, sdd_thresh, percentile)
    fa = _lib.arg_in(f)
    nx, ny = int(fa.obj.shape[0]), int(fa.obj.shape[1])
    r_pts, r_vec, sdd = _host(r_pts), _host(r_vec), _host(sdd)
    cap = int(np.count_nonzero(sdd < 0.0))
    ccap = cap // 2 + 1
    linked = np.empty((max(cap, 1), 2), np.float64)
    ridge_len = np.empty((ccap, 2), np.int32)
    endpoints = np.empty((2 * ccap, 3), np.float64)
    tanvecs = np.empty((2 * ccap, 2), np.float64)
    counts = np.zeros(2, np.int64)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    _lib.check(_lib.load().b200cs_link_ridge_pts(vp(r_pts), vp(r_vec), vp(sdd), nx, ny, float(h), float(c),
                                                 float(sdd_thresh), vp(linked), cap, vp(ridge_len), vp(endpoints),
                                                 vp(tanvecs), ccap, vp(counts)))
    n, k = int(counts[0]), int(counts[1])
    if n > cap or k > ccap:
        raise RuntimeError("ridge linking produced more points than there are ridge pixels")
    return linked[:n], ridge_len[:k], endpoints[:2 * k], tanvecs[:2 * k]


def ftle_ordered_ridges(f, eigvec_max, x, y, dist_tol, ep_tan_ang=pi / 4, min_ridge_pts=5, sdd_thresh=0.0,
                        percentile=0, c=1.0):
    """Ordered, connected FTLE ridges -> list of (k_i, 2) arrays (extraction/ridges.py:720-1054).

    Ridge points are linked into curves (_linked_ridge_pts); curves whose end points are within
    dist_tol of each other, with both curve tangents within ep_tan_ang of the connecting segment,
    are joined; results shorter than min_ridge_pts are dropped."""
    linked, ridge_len, endpoints, tanvecs = _linked_ridge_pts(f, eigvec_max, x, y, sdd_thresh, percentile, c)
    k = len(ridge_len)
    if k == 0:
        return []
    out = np.empty((max(len(linked), 1), 2), np.float64)
    offsets = np.zeros(k + 1, np.int64)
    n_out = np.zeros(1, np.int64)
    linked = np.ascontiguousarray(linked)
    ridge_len, endpoints, tanvecs = (np.ascontiguousarray(a) for a in (ridge_len, endpoints, tanvecs))
    vp = lambda a: C.c_void_p(a.ctypes.data)
    _lib.check(_lib.load().b200cs_order_ridges(vp(linked), len(linked), vp(ridge_len), vp(endpoints), vp(tanvecs), k,
                                               float(dist_tol), float(ep_tan_ang), int(min_ridge_pts), vp(out),
                                               vp(offsets), vp(n_out)))
    m = int(n_out[0])
    return [out[int(offsets[r]):int(offsets[r + 1])].copy() for r in range(m)]
