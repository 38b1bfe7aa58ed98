"""CPU oracle for the NumbaCS flow-map + FTLE + LAVD hot path (ctypes front-end).

THIS IS TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``numbacs_b200`` never does.  The arithmetic lives in ``numbacs_oracle.c`` (see its header for
the reference file:line map and for what is pinned / unpinned); this module only mirrors the
reference's Python call signatures so that parity tests read like the reference's own tests:

    get_predefined_flow   /root/reference/src/numbacs/flows.py:1104
    get_interp_arrays_2D  flows.py:9      get_interp_arrays_scalar  flows.py:85
    get_flow_2D           flows.py:121    get_callable_scalar(_linear)  flows.py:387, 601
    flowmap / flowmap_n / flowmap_grid_2D / flowmap_n_grid_2D   integration.py:7, 64, 123, 467
    ftle_grid_2D / lavd_grid_2D           diagnostics.py:21, 272
    flowmap_aux_grid_2D                   integration.py:249
    C_tensor_2D / C_eig_aux_2D / C_eig_2D / ftle_from_eig   diagnostics.py:68, 115, 200, 247
    ftle_ridge_pts / _ftle_ridge_pts_connect                extraction/ridges.py:9, 232
    composite_simpsons                    utils.py:611
"""
import ctypes as C
import os
import subprocess
from math import pi, sqrt

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnumbacs_oracle.so")

KINDS = {"double_gyre": 0, "bickley_jet": 1, "abc": 2, "spline2d": 3}
EXTRAP = {"constant": 0, "linear": 1, "nearest": 2}


def build(force=False):
    """Compile the C restatement with gcc (seconds)."""
    src = os.path.join(_HERE, "numbacs_oracle.c")
    if (force or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_flow_new.restype = C.c_void_p
        L.oracle_flow_new.argtypes = [C.c_int]
        L.oracle_flow_new_spline.restype = C.c_void_p
        L.oracle_flow_new_spline.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.c_double]
        L.oracle_flow_new_linear.restype = C.c_void_p
        L.oracle_flow_new_linear.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.c_double]
        L.oracle_scalar_new.restype = C.c_void_p
        L.oracle_scalar_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_rhs.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_eval_spline3.restype = C.c_double
        L.oracle_eval_spline3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.oracle_eval_linear3.restype = C.c_double
        L.oracle_eval_linear3.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.oracle_prefilter3.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
        L.oracle_flowmap_pts.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_int64,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p]
        L.oracle_flowmap_grid_2d.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                             C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                             C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
        L.oracle_ftle_grid_2d.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                          C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_composite_simpsons.restype = C.c_double
        L.oracle_composite_simpsons.argtypes = [C.c_void_p, C.c_int64, C.c_double]
        L.oracle_lavd_grid_2d.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                          C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_flowmap_aux_grid_2d.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                                 C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                                 C.c_double, C.c_int, C.c_int, C.c_double,
                                                 C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]
        L.oracle_c_tensor_2d.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_double,
                                         C.c_void_p, C.c_void_p]
        L.oracle_eigh2.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_c_eig_2d.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_c_eig_aux_2d.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_double,
                                          C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p]
        L.oracle_ftle_from_eig.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
        L.oracle_ftle_ridge_pts.restype = C.c_int64
        L.oracle_ftle_ridge_pts.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                            C.c_void_p, C.c_double, C.c_double, C.c_double,
                                            C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_flowmap_composition.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.oracle_scalar_eval_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


class Flow:
    """Owns a C flow_t; keeps borrowed coefficient arrays alive."""

    def __init__(self, handle, ndim, keep=()):
        self.handle = handle
        self.ndim = ndim
        self._keep = keep

    def __del__(self):
        try:
            lib().oracle_free(self.handle)
        except Exception:
            pass

    def rhs(self, t, y, p):
        y = _f64(y)
        p = _f64(p)
        dy = np.zeros(self.ndim)
        lib().oracle_rhs(self.handle, float(t), _ptr(y), _ptr(dy), _ptr(p))
        return dy


class Scalar:
    def __init__(self, handle, linear, keep=()):
        self.handle = handle
        self.linear = linear
        self._keep = keep

    def __del__(self):
        try:
            lib().oracle_free(self.handle)
        except Exception:
            pass

    def __call__(self, pts):
        pts = np.ascontiguousarray(np.atleast_2d(_f64(pts)))
        out = np.empty(len(pts))
        lib().oracle_scalar_eval_many(self.handle, int(self.linear), _ptr(pts), len(pts), _ptr(out))
        return out


def default_params(flow_str, int_direction=1.0):
    """Default parameter vectors and domains of get_predefined_flow (flows.py:1160-1172,
    1215-1238, 1262-1273).  Note the reference forces int_direction=1 for bickley (1219)."""
    if flow_str == "double_gyre":
        return (np.array([int_direction, 0.1, 0.25, 0.0, 0.2 * pi, 0.0]), ((0.0, 2.0), (0.0, 1.0)))
    if flow_str == "bickley_jet":
        r_e = 6371.0e-3
        U0 = 86400 * 62.66e-6
        L = 1770.0e-3
        k1, k2, k3 = 2.0 / r_e, 4.0 / r_e, 6.0 / r_e
        c2 = 0.205 * U0
        c3 = 0.461 * U0
        c1 = c3 + (sqrt(5) - 1) * (c2 - c3)
        return (np.array([1.0, U0, L, 0.0075, 0.15, 0.3, k1, k2, k3, c1, c2, c3]),
                ((0.0, r_e * pi), (-3.0, 3.0)))
    if flow_str == "abc":
        return (np.array([int_direction, 3 ** 0.5, 2 ** 0.5, 1.0, 0.5]),
                ((0.0, 2 * pi), (0.0, 2 * pi), (0.0, 2 * pi)))
    raise ValueError(flow_str)


def get_predefined_flow(flow_str, int_direction=1.0, return_default_params=True,
                        return_domain=True):
    f = Flow(lib().oracle_flow_new(KINDS[flow_str]), 3 if flow_str == "abc" else 2)
    p, dom = default_params(flow_str, int_direction)
    out = [f]
    if return_default_params:
        out.append(p)
    if return_domain:
        out.append(dom)
    return out[0] if len(out) == 1 else tuple(out)


def prefilter(data):
    data = _f64(data)
    n0, n1, n2 = data.shape
    out = np.zeros((n0 + 2, n1 + 2, n2 + 2))
    lib().oracle_prefilter3(_ptr(data), n0, n1, n2, _ptr(out))
    return out


def get_interp_arrays_2D(tvals, xvals, yvals, U, V):
    nt, nx, ny = U.shape
    grid = ((tvals[0], tvals[-1], nt), (xvals[0], xvals[-1], nx), (yvals[0], yvals[-1], ny))
    return grid, prefilter(U), prefilter(V)


def get_interp_arrays_scalar(tvals, xvals, yvals, f):
    nt, nx, ny = f.shape
    if tvals[1] < tvals[0]:
        f = np.flip(f, axis=0)
        tvals = tvals[::-1]
    grid = ((tvals[0], tvals[-1], nt), (xvals[0], xvals[-1], nx), (yvals[0], yvals[-1], ny))
    return grid, prefilter(f)


def _grid9(grid):
    return _f64(np.array([[g[0], g[1], float(g[2])] for g in grid]).ravel())


def get_flow_2D(grid_vel, C_eval_u, C_eval_v, spherical=0, extrap_mode="constant", r=6371.0):
    g = _grid9(grid_vel)
    cu, cv = _f64(C_eval_u), _f64(C_eval_v)
    h = lib().oracle_flow_new_spline(_ptr(g), _ptr(cu), _ptr(cv), int(spherical),
                                     EXTRAP[extrap_mode], float(r))
    return Flow(h, 2, keep=(g, cu, cv))


def get_flow_linear_2D(grid_vel, U, V, spherical=0, extrap_mode="constant", r=6371.0):
    g = _grid9(grid_vel)
    u, v = _f64(U), _f64(V)
    h = lib().oracle_flow_new_linear(_ptr(g), _ptr(u), _ptr(v), int(spherical), EXTRAP[extrap_mode],
                                     float(r))
    return Flow(h, 2, keep=(g, u, v))


def get_callable_scalar(grid_f, C_eval_f, extrap_mode="constant"):
    g = _grid9(grid_f)
    c = _f64(C_eval_f)
    return Scalar(lib().oracle_scalar_new(_ptr(g), _ptr(c), EXTRAP[extrap_mode]), False, (g, c))


def get_callable_scalar_linear(grid_f, f, extrap_mode="constant"):
    g = _grid9(grid_f)
    c = _f64(f)
    return Scalar(lib().oracle_scalar_new(_ptr(g), _ptr(c), EXTRAP[extrap_mode]), True, (g, c))


def _mask(mask):
    if mask is None:
        return None
    return np.ascontiguousarray(mask, dtype=np.uint8)


def flowmap_pts(flow, t0, T, pts, params, n=2, last_only=True, rtol=1e-6, atol=1e-8, mask=None,
                full=False):
    pts = _f64(pts)
    params = _f64(params)
    npts, nd = pts.shape
    assert nd == flow.ndim
    out = np.zeros((npts, nd) if last_only else (npts, n, nd))
    tspan = np.zeros(n)
    status = np.zeros(npts, np.int32)
    steps = np.zeros((npts, 2), np.int32)
    stats = np.zeros(3, np.int64)
    m = _mask(mask)
    lib().oracle_flowmap_pts(flow.handle, float(t0), float(T), _ptr(pts), npts, _ptr(params),
                             float(rtol), float(atol), _ptr(m), n, int(last_only), _ptr(out),
                             _ptr(tspan), _ptr(status), _ptr(steps), _ptr(stats))
    if full:
        return out, tspan, status, steps, stats
    return out if last_only else (out, tspan)


def flowmap(flow, t0, T, pts, params, method="dop853", rtol=1e-6, atol=1e-8, mask=None):
    return flowmap_pts(flow, t0, T, pts, params, 2, True, rtol, atol, mask)


def flowmap_n(flow, t0, T, pts, params, method="dop853", n=2, rtol=1e-6, atol=1e-8, mask=None):
    return flowmap_pts(flow, t0, T, pts, params, n, False, rtol, atol, mask)


def _grid(flow, t0, T, x, y, params, n, last_only, rtol, atol, mask, full):
    x, y, params = _f64(x), _f64(y), _f64(params)
    nx, ny = len(x), len(y)
    out = np.zeros((nx, ny, 2) if last_only else (nx, ny, n, 2))
    tspan = np.zeros(n)
    status = np.zeros((nx, ny), np.int32)
    steps = np.zeros((nx, ny, 2), np.int32)
    stats = np.zeros(3, np.int64)
    m = _mask(mask)
    lib().oracle_flowmap_grid_2d(flow.handle, float(t0), float(T), _ptr(x), nx, _ptr(y), ny,
                                 _ptr(params), float(rtol), float(atol), _ptr(m), n,
                                 int(last_only), _ptr(out), _ptr(tspan), _ptr(status), _ptr(steps),
                                 _ptr(stats))
    if full:
        return out, tspan, status, steps, stats
    return out if last_only else (out, tspan)


def flowmap_grid_2D(flow, t0, T, x, y, params, method="dop853", rtol=1e-6, atol=1e-8, mask=None,
                    full=False):
    return _grid(flow, t0, T, x, y, params, 2, True, rtol, atol, mask, full)


def flowmap_n_grid_2D(flow, t0, T, x, y, params, n=50, method="dop853", rtol=1e-6, atol=1e-8,
                      mask=None, full=False):
    return _grid(flow, t0, T, x, y, params, n, False, rtol, atol, mask, full)


def ftle_grid_2D(flowmap, T, dx, dy, mask=None):
    fm = _f64(flowmap)
    nx, ny = fm.shape[:2]
    out = np.zeros((nx, ny))
    m = _mask(mask)
    lib().oracle_ftle_grid_2d(_ptr(fm), nx, ny, float(T), float(dx), float(dy), _ptr(m), _ptr(out))
    return out


def composite_simpsons(f, h):
    f = _f64(f)
    return lib().oracle_composite_simpsons(_ptr(f), len(f), float(h))


def lavd_grid_2D(flowmap_n, tspan, T, vort_interp, xrav, yrav, period_x=0.0, period_y=0.0,
                 mask=None):
    fm = _f64(flowmap_n)
    nx, ny, n = fm.shape[:3]
    tspan, xrav, yrav = _f64(tspan), _f64(xrav), _f64(yrav)
    out = np.zeros((nx, ny))
    m = _mask(mask)
    lib().oracle_lavd_grid_2d(_ptr(fm), nx, ny, n, _ptr(tspan), vort_interp.handle,
                              int(vort_interp.linear), _ptr(xrav), _ptr(yrav), float(period_x),
                              float(period_y), _ptr(m), _ptr(out))
    return out


def flowmap_aux_grid_2D(flow, t0, T, x, y, params, h=1e-5, eig_main=True, compute_edge=True,
                        method="dop853", rtol=1e-6, atol=1e-8, mask=None, full=False):
    x, y, params = _f64(x), _f64(y), _f64(params)
    nx, ny = len(x), len(y)
    n_aux = 5 if eig_main else 4
    out = np.zeros((nx, ny, n_aux, 2))
    status = np.zeros((nx, ny, n_aux), np.int32)
    steps = np.zeros((nx, ny, n_aux, 2), np.int32)
    stats = np.zeros(3, np.int64)
    m = _mask(mask)
    lib().oracle_flowmap_aux_grid_2d(flow.handle, float(t0), float(T), _ptr(x), nx, _ptr(y), ny,
                                     _ptr(params), float(h), int(eig_main), int(compute_edge),
                                     float(rtol), float(atol), _ptr(m), _ptr(out), _ptr(status),
                                     _ptr(steps), _ptr(stats))
    return (out, status, steps, stats) if full else out


def C_tensor_2D(flowmap_aux, dx, dy, h=1e-5, mask=None):
    fa = _f64(flowmap_aux)
    nx, ny, n_aux = fa.shape[:3]
    out = np.zeros((nx, ny, 3))
    m = _mask(mask)
    lib().oracle_c_tensor_2d(_ptr(fa), nx, ny, n_aux, float(h), _ptr(m), _ptr(out))
    return out


def eigh2(a, b, c):
    w, v = np.zeros(2), np.zeros((2, 2))
    lib().oracle_eigh2(float(a), float(b), float(c), _ptr(w), _ptr(v))
    return w, v


def C_eig_2D(flowmap, dx, dy, mask=None):
    fm = _f64(flowmap)
    nx, ny = fm.shape[:2]
    vals, vecs = np.zeros((nx, ny, 2)), np.zeros((nx, ny, 2, 2))
    m = _mask(mask)
    lib().oracle_c_eig_2d(_ptr(fm), nx, ny, float(dx), float(dy), _ptr(m), _ptr(vals), _ptr(vecs))
    return vals, vecs


def C_eig_aux_2D(flowmap_aux, dx, dy, h=1e-5, eig_main=True, mask=None):
    fa = _f64(flowmap_aux)
    nx, ny, n_aux = fa.shape[:3]
    vals, vecs = np.zeros((nx, ny, 2)), np.zeros((nx, ny, 2, 2))
    m = _mask(mask)
    lib().oracle_c_eig_aux_2d(_ptr(fa), nx, ny, n_aux, float(dx), float(dy), float(h),
                              int(eig_main), _ptr(m), _ptr(vals), _ptr(vecs))
    return vals, vecs


def ftle_from_eig(eigval_max, T):
    e = _f64(eigval_max)
    out = np.zeros(e.shape)
    lib().oracle_ftle_from_eig(_ptr(e), e.size, float(T), _ptr(out))
    return out


def _ftle_ridge_pts_connect(f, eigvec_max, x, y, sdd_thresh=0.0, percentile=0, spacing=None):
    f, ev, x, y = _f64(f), _f64(eigvec_max), _f64(x), _f64(y)
    dx, dy = (x[1] - x[0], y[1] - y[0]) if spacing is None else spacing
    nx, ny = f.shape
    f_min = 0.0 if percentile == 0 else float(np.percentile(f, percentile))
    r_pts, r_vec, sdd = np.zeros((nx * ny, 3)), np.zeros((nx * ny, 2)), np.zeros(nx * ny)
    lib().oracle_ftle_ridge_pts(_ptr(f), _ptr(ev), nx, ny, _ptr(x), _ptr(y), float(dx), float(dy),
                                float(sdd_thresh), f_min, _ptr(r_pts), _ptr(r_vec), _ptr(sdd))
    return r_pts, r_vec, sdd, min(x[1] - x[0], y[1] - y[0])


def ftle_ridge_pts(f, eigvec_max, x, y, sdd_thresh=0.0, percentile=0, spacing=None):
    r_pts, _, sdd, _ = _ftle_ridge_pts_connect(f, eigvec_max, x, y, sdd_thresh, percentile, spacing)
    return r_pts[sdd < -sdd_thresh][:, :2]  # sdd is c2 (< -sdd_thresh <= 0) at ridge points, else 0


def ftle_ridges(f, eigvec_max, x, y, sdd_thresh=0.0, percentile=0, min_ridge_pts=3):
    """ftle_ridges (extraction/ridges.py:175-229): the reference labels ridge_bool with
    scipy.ndimage.label (8-connectivity) -- third-party code present here, called the same way --
    and lists each label's points in raveled order."""
    from scipy.ndimage import generate_binary_structure, label
    f = _f64(f)
    nx, ny = f.shape
    r_pts, _, sdd, _ = _ftle_ridge_pts_connect(f, eigvec_max, x, y, sdd_thresh, percentile)
    ridge_bool = (sdd < -sdd_thresh).reshape(nx, ny)
    labels, nlabels = label(ridge_bool, structure=generate_binary_structure(2, 2))
    flat = labels.ravel()
    out = []
    for i in range(1, nlabels + 1):
        inds = np.flatnonzero(flat == i)
        if len(inds) >= min_ridge_pts:
            out.append(r_pts[inds, :2])
    return out


def _grid6(grid):
    return _f64(np.array([[g[0], g[1], float(g[2])] for g in grid]).ravel())


def flowmap_composition(flowmaps, grid, nT):
    fms = _f64(flowmaps)
    g = _grid6(grid)
    nx, ny = int(grid[0][2]), int(grid[1][2])
    out = np.zeros((nx, ny, 2))
    lib().oracle_flowmap_composition(_ptr(fms), _ptr(g), int(nT), _ptr(out))
    return out


def flowmap_composition_initial(flow, t0, T, h, x, y, grid, params, **kwargs):
    nT = abs(round(T / h))
    flowmaps = np.zeros((nT, int(grid[0][2]), int(grid[1][2]), 2))
    for k in range(nT):
        flowmaps[k] = flowmap_grid_2D(flow, t0, h, x, y, params, **kwargs)
        t0 += h
    return flowmap_composition(flowmaps, grid, nT), flowmaps, nT


def flowmap_composition_step(flowmaps, flow, t0, h, nT, x, y, grid, params, **kwargs):
    flowmaps = _f64(flowmaps).copy()
    flowmaps[:-1] = flowmaps[1:].copy()
    flowmaps[-1] = flowmap_grid_2D(flow, t0, h, x, y, params, **kwargs)
    return flowmap_composition(flowmaps, grid, nT), flowmaps


def binary_mask_dilation(mask, corners=False):
    """binary_mask_dilation (utils.py:1923-1985), restated with shifted views."""
    m = np.asarray(mask, dtype=bool)
    out = m.copy()
    out[1:] |= m[:-1]
    out[:-1] |= m[1:]
    out[:, 1:] |= m[:, :-1]
    out[:, :-1] |= m[:, 1:]
    if corners:
        out[1:, 1:] |= m[:-1, :-1]
        out[1:, :-1] |= m[:-1, 1:]
        out[:-1, 1:] |= m[1:, :-1]
        out[:-1, :-1] |= m[1:, 1:]
    return out
