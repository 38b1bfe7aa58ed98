"""Flow construction -- drop-in for the hot-path part of ``numbacs.flows``.

Same names, arguments, defaults and return layouts as the reference
(/root/reference/src/numbacs/flows.py): get_predefined_flow (1104), get_interp_arrays_2D (9),
get_interp_arrays_scalar (85), get_flow_2D (121), get_callable_scalar (387),
get_callable_scalar_linear (601).

The one semantic change: ``funcptr`` is still an ``int`` but it is a *handle* into the GPU-side
flow registry of libb200cs.so instead of the address of a numba ``@cfunc`` -- a GPU cannot call
a CPU function pointer.  Handles are only meaningful to ``numbacs_b200.integration``.
"""
import ctypes as C
from math import pi, sqrt

import numpy as np

from . import _lib

__all__ = ["get_predefined_flow", "get_interp_arrays_2D", "get_interp_arrays_scalar", "get_flow_linear_2D",
           "get_flow_2D", "get_callable_2D", "get_callable_scalar", "get_callable_scalar_linear", "ScalarField",
           "release_flow"]


def _grid9(grid):
    g = np.array([[float(ax[0]), float(ax[1]), float(ax[2])] for ax in grid], dtype=np.float64)
    if g.shape != (3, 3):
        raise ValueError("grid must be ((t0,t1,nt),(x0,x1,nx),(y0,y1,ny))")
    return np.ascontiguousarray(g.ravel())


def get_predefined_flow(flow_str, int_direction=1.0, return_default_params=True,
                        return_domain=True, parameter_description=False):
    """Handle of the device implementation of a predefined flow ('double_gyre', 'bickley_jet',
    'abc'), with the reference's default parameters / domain / description (flows.py:1104-1295).
    """
    if flow_str not in _lib.FLOW_KINDS:
        raise ValueError(f"unknown flow {flow_str!r}; supported: {sorted(_lib.FLOW_KINDS)}")
    h = C.c_int(0)
    _lib.check(_lib.load().b200cs_flow_create_analytic(_lib.FLOW_KINDS[flow_str], C.byref(h)))
    funcptr = h.value

    if flow_str == "double_gyre":
        default_params = np.array([int_direction, 0.1, 0.25, 0.0, 0.2 * pi, 0.0])
        domain = ((0.0, 2.0), (0.0, 1.0))
        p_str = ("p[0] = int_direction, p[1] = A, p[2] = eps, p[3] = alpha, "
                 "p[4] = omega, p[5] = psi")
    elif flow_str == "bickley_jet":
        # units: time - days, length - Mm.  The reference forces int_direction = 1 in the default
        # parameters whatever the argument says (flows.py:1219); kept for parity.
        r_e = 6371.0e-3
        U0 = 86400 * 62.66e-6
        L = 1770.0e-3
        k1, k2, k3 = 2.0 / r_e, 4.0 / r_e, 6.0 / r_e
        c2 = 0.205 * U0
        c3 = 0.461 * U0
        c1 = c3 + (sqrt(5) - 1) * (c2 - c3)
        default_params = np.array([1.0, U0, L, 0.0075, 0.15, 0.3, k1, k2, k3, c1, c2, c3])
        domain = ((0.0, r_e * pi), (-3.0, 3.0))
        p_str = ("p[0] = int_direction, p[1] = U0, p[2] = L, p[3] = A1, p[4] = A2, "
                 "p[5] = A3, p[6] = k1, p[7] = k2, p[8] = k3, p[9] = c1, p[10] = c2, "
                 "p[11] = c3, units: time - days, length - Mm")
    else:
        default_params = np.array([int_direction, 3 ** 0.5, 2 ** 0.5, 1.0, 0.5])
        domain = ((0.0, 2 * pi), (0.0, 2 * pi), (0.0, 2 * pi))
        p_str = ("p[0] = int_direction, p[1] = A-amplitude, p[2] = B-amplitude, "
                 "p[3] = C-amplitude, p[4] = forcing amplitdue")

    out = [funcptr]
    if return_default_params:
        out.append(default_params)
    if return_domain:
        out.append(domain)
    if parameter_description:
        out.append(p_str)
    return out[0] if len(out) == 1 else tuple(out)


def _prefilter(f):
    f_in = _lib.arg_in(f)
    shp = tuple(f_in.obj.shape)
    if len(shp) != 3:
        raise ValueError("expected a (nt, nx, ny) array")
    dev = f_in.on_device
    out = _lib.alloc_out((shp[0] + 2, shp[1] + 2, shp[2] + 2), np.float64, dev)
    _lib.check(_lib.load().b200cs_prefilter_3d(f_in.ptr, shp[0], shp[1], shp[2], out.ptr,
                                               _lib.current_stream(dev)))
    return out.obj


def get_interp_arrays_2D(tvals, xvals, yvals, U, V):
    """Cubic B-spline coefficient arrays of the velocity field (flows.py:9-46); the natural-BC
    prefilter runs on the GPU.  Returns (grid_vel, C_eval_u, C_eval_v), arrays (nt+2,nx+2,ny+2)."""
    nt, nx, ny = U.shape
    grid_vel = ((float(tvals[0]), float(tvals[-1]), nt), (float(xvals[0]), float(xvals[-1]), nx),
                (float(yvals[0]), float(yvals[-1]), ny))
    return grid_vel, _prefilter(U), _prefilter(V)


def get_interp_arrays_scalar(tvals, xvals, yvals, f):
    """Cubic B-spline coefficients of a scalar field (flows.py:85-118); descending `tvals`
    are flipped like the reference does."""
    nt, nx, ny = f.shape
    if tvals[1] < tvals[0]:
        f = f.flip(0) if _lib._is_torch(f) else np.flip(f, axis=0)
        tvals = tvals[::-1] if not _lib._is_torch(tvals) else tvals.flip(0)
    grid_f = ((float(tvals[0]), float(tvals[-1]), nt), (float(xvals[0]), float(xvals[-1]), nx),
              (float(yvals[0]), float(yvals[-1]), ny))
    return grid_f, _prefilter(f)


def get_flow_2D(grid_vel, C_eval_u, C_eval_v, spherical=0, extrap_mode="constant", r=6371.0):
    """Handle of the device cubic-spline flow over (grid_vel, C_eval_u, C_eval_v)
    (flows.py:121-258): spherical 0 / 1 (lon in [-180,180)) / 2 (lon in [0,360)), degrees."""
    if extrap_mode not in _lib.EXTRAP:
        raise ValueError(f"unknown extrap_mode {extrap_mode!r}")
    g = _grid9(grid_vel)
    cu, cv = _lib.arg_in(C_eval_u), _lib.arg_in(C_eval_v)
    exp = (int(g[2]) + 2, int(g[5]) + 2, int(g[8]) + 2)
    if tuple(cu.obj.shape) != exp or tuple(cv.obj.shape) != exp:
        raise ValueError(f"coefficient arrays must have shape {exp}")
    h = C.c_int(0)
    _lib.check(_lib.load().b200cs_flow_create_spline(
        C.c_void_p(g.ctypes.data), cu.ptr, cv.ptr, int(spherical), _lib.EXTRAP[extrap_mode],
        float(r), C.byref(h)))
    return h.value


def get_flow_linear_2D(grid_vel, U, V, spherical=0, extrap_mode="constant", r=6371.0):
    """Handle of the device trilinear flow over the raw velocity arrays U, V (nt, nx, ny)
    (flows.py:418-506); same conventions as get_flow_2D."""
    if extrap_mode not in _lib.EXTRAP:
        raise ValueError(f"unknown extrap_mode {extrap_mode!r}")
    g = _grid9(grid_vel)
    ua, va = _lib.arg_in(U), _lib.arg_in(V)
    exp = (int(g[2]), int(g[5]), int(g[8]))
    if tuple(ua.obj.shape) != exp or tuple(va.obj.shape) != exp:
        raise ValueError(f"velocity arrays must have shape {exp}")
    h = C.c_int(0)
    _lib.check(_lib.load().b200cs_flow_create_linear(
        C.c_void_p(g.ctypes.data), ua.ptr, va.ptr, int(spherical), _lib.EXTRAP[extrap_mode],
        float(r), C.byref(h)))
    return h.value


class VelocityField:
    """Device-resident velocity interpolant (the object get_callable_2D returns).  Callable like the
    reference's jit function: vel(point[3]) -> array([u, v]) (or the tuple (u, v) with
    return_type="tuple"); vel(points[N, 3]) -> array[N, 2].  numbacs_b200.utils.curl_func_tspan
    recognises it and differentiates it on the GPU."""

    def __init__(self, handle, return_type):
        self.handle = handle
        self.return_type = return_type

    def __call__(self, pts):
        single = (not _lib._is_torch(pts)) and np.ndim(pts) == 1
        p = _lib.arg_in(np.atleast_2d(pts) if single else pts)
        n = int(p.obj.shape[0])
        out = _lib.alloc_out((n, 2), np.float64, p.on_device)
        _lib.check(_lib.load().b200cs_velocity_eval(self.handle, p.ptr, n, out.ptr,
                                                    _lib.current_stream(p.on_device)))
        if not single:
            return out.obj
        if self.return_type == "tuple":
            return float(out.obj[0, 0]), float(out.obj[0, 1])
        return np.array([out.obj[0, 0], out.obj[0, 1]], np.float64)


def get_callable_2D(grid_vel, C_eval_u, C_eval_v, spherical=0, extrap_mode="constant", r=6371.0,
                    return_type="array"):
    """Callable spline of the velocity field (flows.py:261-384), evaluated on the device.  As in the
    reference only spherical == 1 applies the spherical scaling; there is no longitude wrap and no
    params[0] in a callable."""
    if return_type not in ("array", "tuple"):
        raise ValueError("return_type must be 'array' or 'tuple'")
    h = get_flow_2D(grid_vel, C_eval_u, C_eval_v, 1 if spherical == 1 else 0, extrap_mode, r)
    return VelocityField(h, return_type)


def release_flow(funcptr):
    """Free the device memory held by a flow / scalar handle (optional; handles otherwise live
    for the life of the process, like the reference's cfuncs)."""
    _lib.check(_lib.load().b200cs_flow_destroy(int(funcptr)))


class ScalarField:
    """Device-resident scalar interpolant (the object get_callable_scalar* return).  Callable
    like the reference's jit functions: f(point[3]) -> float, f(points[N,3]) -> array[N]."""

    def __init__(self, handle, linear):
        self.handle = handle
        self.linear = linear

    def __call__(self, pts):
        single = (not _lib._is_torch(pts)) and np.ndim(pts) == 1
        p = _lib.arg_in(np.atleast_2d(pts) if single else pts)
        n = int(p.obj.shape[0])
        out = _lib.alloc_out((n,), np.float64, p.on_device)
        _lib.check(_lib.load().b200cs_scalar_eval(self.handle, p.ptr, n, out.ptr,
                                                  _lib.current_stream(p.on_device)))
        return float(out.obj[0]) if single else out.obj


def _scalar(grid_f, data, linear, extrap_mode):
    if extrap_mode not in _lib.EXTRAP:
        raise ValueError(f"unknown extrap_mode {extrap_mode!r}")
    g = _grid9(grid_f)
    d = _lib.arg_in(data)
    pad = 0 if linear else 2
    exp = (int(g[2]) + pad, int(g[5]) + pad, int(g[8]) + pad)
    if tuple(d.obj.shape) != exp:
        raise ValueError(f"array must have shape {exp}")
    h = C.c_int(0)
    _lib.check(_lib.load().b200cs_scalar_create(C.c_void_p(g.ctypes.data), d.ptr, int(linear),
                                                _lib.EXTRAP[extrap_mode], C.byref(h)))
    return ScalarField(h.value, bool(linear))


def get_callable_scalar(grid_f, C_eval_f, extrap_mode="constant"):
    """Cubic-spline scalar interpolant on the device (flows.py:387-415)."""
    return _scalar(grid_f, C_eval_f, False, extrap_mode)


def get_callable_scalar_linear(grid_f, f, extrap_mode="constant"):
    """Trilinear scalar interpolant on the device (flows.py:601-636)."""
    return _scalar(grid_f, f, True, extrap_mode)
