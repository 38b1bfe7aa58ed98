"""Particle integration -- drop-in for the hot-path part of ``numbacs.integration``.

flowmap (integration.py:7), flowmap_n (64), flowmap_grid_2D (123), flowmap_aux_grid_2D (249),
flowmap_n_grid_2D (467) of the reference, same positional arguments, defaults (method="dop853", rtol=1e-6, atol=1e-8, mask=None)
and output layouts ('ij' indexing, float64, zeros where masked).  The work is done by the
one-thread-per-particle DOP853 kernels of libb200cs.so; ``funcptr`` must be a handle from
``numbacs_b200.flows``.

Extras (keyword-only, not in the reference):
  device_out : return torch CUDA tensors instead of numpy arrays (no PCIe round trip before
               ftle_grid_2D / lavd_grid_2D, which accept them).  Implied when an input is a
               CUDA tensor.
  info       : a dict that receives 'status' (per-particle numbalsoda `success`, which the
               reference silently drops), 'steps' (accepted, rejected attempts per particle) and
               'stats' = [sum nfev, sum accepted, sum rejected].
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["flowmap", "flowmap_n", "flowmap_grid_2D", "flowmap_n_grid_2D", "flowmap_aux_grid_2D",
           "flowmap_grid_ND", "flowmap_n_grid_ND", "flowmap_grid_2D_series",
           "flowmap_composition", "flowmap_composition_initial", "flowmap_composition_step",
           "flowmap_composition_series"]


def _method(method):
    m = method.lower()
    if m == "dop853":
        return _lib.METHOD_DOP853
    if m == "lsoda":
        raise NotImplementedError("method='lsoda' is not implemented on the GPU; use 'dop853'")
    raise ValueError(f"unknown method {method!r}")


def _info_bufs(info, shape, device):
    if info is None:
        return _lib.Arg(None, None, False), _lib.Arg(None, None, False), _lib.Arg(None, None, False)
    status = _lib.alloc_out(shape, np.int32, device)
    steps = _lib.alloc_out(shape + (2,), np.int32, device)
    if device:
        import torch
        st = torch.zeros(3, dtype=torch.int64, device="cuda")
        stats = _lib.Arg(st, C.c_void_p(st.data_ptr()), True)
    else:
        st = np.zeros(3, np.int64)
        stats = _lib.Arg(st, C.c_void_p(st.ctypes.data), False)
    return status, steps, stats


def out_of_grid_count(funcptr, reset=True, device=False):
    """Evaluations of an interpolated flow outside its data grid since the last reset (0 for analytic
    flows): the guard on the unpinned extrapolation modes (b200cs_flow_out_of_grid)."""
    n = np.zeros(1, np.int64)
    _lib.check(_lib.load().b200cs_flow_out_of_grid(int(funcptr), C.c_void_p(n.ctypes.data), int(bool(reset)),
                                                   _lib.current_stream(device)))
    return int(n[0])


def _fill_info(info, status, steps, stats):
    if info is not None:
        info["status"], info["steps"], info["stats"] = status.obj, steps.obj, stats.obj


def _grid(funcptr, t0, T, x, y, params, n, method, rtol, atol, mask, device_out, info):
    xa, ya, pa, ma = _lib.arg_in(x), _lib.arg_in(y), _lib.arg_in(params), _lib.mask_in(mask)
    nx, ny = int(xa.obj.shape[0]), int(ya.obj.shape[0])
    if ma.obj is not None and tuple(ma.obj.shape) != (nx, ny):
        raise ValueError(f"mask must have shape {(nx, ny)}")
    dev = bool(device_out or xa.on_device or ya.on_device or ma.on_device)
    out = _lib.alloc_out((nx, ny, 2) if n == 0 else (nx, ny, n, 2), np.float64, dev)
    tspan = np.empty(max(n, 1), np.float64)
    status, steps, stats = _info_bufs(info, (nx, ny), dev)
    _lib.check(_lib.load().b200cs_flowmap_grid_2d(
        int(funcptr), float(t0), float(T), xa.ptr, nx, ya.ptr, ny, pa.ptr, int(pa.obj.shape[0]),
        _method(method), float(rtol), float(atol), ma.ptr, int(n), out.ptr,
        C.c_void_p(tspan.ctypes.data), status.ptr, steps.ptr, stats.ptr, _lib.current_stream(dev)))
    _fill_info(info, status, steps, stats)
    if info is not None:
        info["out_of_grid"] = out_of_grid_count(funcptr, True, dev)
    return out.obj, tspan


def _pts(funcptr, t0, T, pts, params, n, method, rtol, atol, mask, device_out, info):
    qa, pa, ma = _lib.arg_in(pts), _lib.arg_in(params), _lib.mask_in(mask)
    if qa.obj.ndim != 2:
        raise ValueError("pts must have shape (npts, N)")
    npts, nd = int(qa.obj.shape[0]), int(qa.obj.shape[1])
    dev = bool(device_out or qa.on_device or ma.on_device)
    out = _lib.alloc_out((npts, nd) if n == 0 else (npts, n, nd), np.float64, dev)
    tspan = np.empty(max(n, 1), np.float64)
    status, steps, stats = _info_bufs(info, (npts,), dev)
    _lib.check(_lib.load().b200cs_flowmap_pts(
        int(funcptr), float(t0), float(T), qa.ptr, npts, nd, pa.ptr, int(pa.obj.shape[0]),
        _method(method), float(rtol), float(atol), ma.ptr, int(n), out.ptr,
        C.c_void_p(tspan.ctypes.data), status.ptr, steps.ptr, stats.ptr, _lib.current_stream(dev)))
    _fill_info(info, status, steps, stats)
    if info is not None:
        info["out_of_grid"] = out_of_grid_count(funcptr, True, dev)
    return out.obj, tspan


def flowmap(funcptr, t0, T, pts, params, method="dop853", rtol=1e-6, atol=1e-8, mask=None, *,
            device_out=False, info=None):
    """Final positions of the particles pts[npts, N] after [t0, t0+T] -> (npts, N)."""
    return _pts(funcptr, t0, T, pts, params, 0, method, rtol, atol, mask, device_out, info)[0]


def flowmap_n(funcptr, t0, T, pts, params, method="dop853", n=2, rtol=1e-6, atol=1e-8, mask=None,
              *, device_out=False, info=None):
    """Positions at n equally spaced times -> ((npts, n, N), t_eval[n])."""
    return _pts(funcptr, t0, T, pts, params, n, method, rtol, atol, mask, device_out, info)


def flowmap_grid_2D(funcptr, t0, T, x, y, params, method="dop853", rtol=1e-6, atol=1e-8,
                    mask=None, *, device_out=False, info=None):
    """Flow map at the final time over the 'ij' grid (x, y) -> (nx, ny, 2)."""
    return _grid(funcptr, t0, T, x, y, params, 0, method, rtol, atol, mask, device_out, info)[0]


def flowmap_n_grid_2D(funcptr, t0, T, x, y, params, n=50, method="dop853", rtol=1e-6, atol=1e-8,
                      mask=None, *, device_out=False, info=None):
    """Flow map at n equally spaced times over the grid -> ((nx, ny, n, 2), t_eval[n])."""
    return _grid(funcptr, t0, T, x, y, params, n, method, rtol, atol, mask, device_out, info)


def flowmap_aux_grid_2D(funcptr, t0, T, x, y, params, h=1e-5, eig_main=True, compute_edge=True,
                        method="dop853", rtol=1e-6, atol=1e-8, mask=None, *, device_out=False,
                        info=None):
    """Flow map at the final time over the auxiliary grid (x, y) +- h -> (nx, ny, n_aux, 2) with
    n_aux = 5 (eig_main: the grid point itself is the last entry) or 4.  Entries the reference
    leaves untouched (masked cells, the stencil points of edge cells) are 0."""
    xa, ya, pa, ma = _lib.arg_in(x), _lib.arg_in(y), _lib.arg_in(params), _lib.mask_in(mask)
    nx, ny = int(xa.obj.shape[0]), int(ya.obj.shape[0])
    if ma.obj is not None and tuple(ma.obj.shape) != (nx, ny):
        raise ValueError(f"mask must have shape {(nx, ny)}")
    n_aux = 5 if eig_main else 4
    dev = bool(device_out or xa.on_device or ya.on_device or ma.on_device)
    out = _lib.alloc_out((nx, ny, n_aux, 2), np.float64, dev)
    status, steps, stats = _info_bufs(info, (nx, ny, n_aux), dev)
    _lib.check(_lib.load().b200cs_flowmap_aux_grid_2d(
        int(funcptr), float(t0), float(T), xa.ptr, nx, ya.ptr, ny, pa.ptr, int(pa.obj.shape[0]),
        float(h), int(bool(eig_main)), int(bool(compute_edge)), _method(method), float(rtol),
        float(atol), ma.ptr, out.ptr, status.ptr, steps.ptr, stats.ptr, _lib.current_stream(dev)))
    _fill_info(info, status, steps, stats)
    return out.obj


def flowmap_grid_2D_series(funcptr, t0s, T, x, y, params, method="dop853", rtol=1e-6, atol=1e-8,
                           mask=None, *, device_out=False, info=None, out=None):
    """flowmap_grid_2D for a whole series of initial times in ONE launch -> (nt, nx, ny, 2).

    Frame f is bit-identical to flowmap_grid_2D(funcptr, t0s[f], T, x, y, params, ...).  An FTLE
    movie frame (201 x 101 in the reference's time-series examples) is far too small to fill the
    GPU; the batch is what does.  `out` (optional) is a preallocated (nt, nx, ny, 2) float64 numpy
    array or CUDA tensor to write into."""
    xa, ya, pa, ma = _lib.arg_in(x), _lib.arg_in(y), _lib.arg_in(params), _lib.mask_in(mask)
    ta = _lib.arg_in(np.atleast_1d(t0s) if not _lib._is_torch(t0s) else t0s)
    nt, nx, ny = int(ta.obj.shape[0]), int(xa.obj.shape[0]), int(ya.obj.shape[0])
    if ma.obj is not None and tuple(ma.obj.shape) != (nx, ny):
        raise ValueError(f"mask must have shape {(nx, ny)}")
    if out is not None:
        if tuple(int(v) for v in out.shape) != (nt, nx, ny, 2):
            raise ValueError(f"out must have shape {(nt, nx, ny, 2)}")
        on_dev = _lib._is_torch(out)
        if on_dev and not (out.is_cuda and out.is_contiguous() and str(out.dtype) == "torch.float64"):
            raise ValueError("out must be a contiguous float64 CUDA tensor")
        if not on_dev and not (out.dtype == np.float64 and out.flags.c_contiguous):
            raise ValueError("out must be a C-contiguous float64 array")
        res = _lib.Arg(out, C.c_void_p(out.data_ptr() if on_dev else out.ctypes.data), on_dev)
        dev = on_dev
    else:
        dev = bool(device_out or xa.on_device or ya.on_device or ma.on_device or ta.on_device)
        res = _lib.alloc_out((nt, nx, ny, 2), np.float64, dev)
    status, steps, stats = _info_bufs(info, (nt, nx, ny), dev)
    _lib.check(_lib.load().b200cs_flowmap_grid_2d_series(
        int(funcptr), ta.ptr, nt, float(T), xa.ptr, nx, ya.ptr, ny, pa.ptr, int(pa.obj.shape[0]),
        _method(method), float(rtol), float(atol), ma.ptr, res.ptr, status.ptr, steps.ptr, stats.ptr,
        _lib.current_stream(dev)))
    _fill_info(info, status, steps, stats)
    return res.obj


def _flat_points(IC_flat, ndims):
    """IC_flat (npts * ndims,) -> (npts, ndims) view: particle k is IC_flat[k*ndims:(k+1)*ndims]
    (integration.py:224, 586), i.e. exactly the row-major point list the pts kernels take."""
    nd = int(ndims)
    if _lib._is_torch(IC_flat):
        flat = IC_flat.reshape(-1)
    else:
        flat = np.asarray(IC_flat, dtype=np.float64).reshape(-1)
    npts = int(flat.shape[0] // nd)
    return flat[:npts * nd].reshape(npts, nd)


def flowmap_grid_ND(funcptr, t0, T, IC_flat, ndims, params, method="dop853", rtol=1e-6, atol=1e-8,
                    *, device_out=False, info=None):
    """Final positions for a flattened list of ndims-dimensional initial conditions -> (npts, ndims)
    (integration.py:185-246; e.g. a 3-D grid for the abc flow)."""
    return _pts(funcptr, t0, T, _flat_points(IC_flat, ndims), params, 0, method, rtol, atol, None,
                device_out, info)[0]


def flowmap_n_grid_ND(funcptr, t0, T, IC_flat, ndims, params, n=50, method="dop853", rtol=1e-6,
                      atol=1e-8, *, device_out=False, info=None):
    """Positions at n equally spaced times -> ((npts, n, ndims), t_eval[n]) (integration.py:536-606)."""
    return _pts(funcptr, t0, T, _flat_points(IC_flat, ndims), params, n, method, rtol, atol, None,
                device_out, info)


# ---- flow-map composition (integration.py:609-737): FTLE time series from short flow maps ----------

def _grid6(grid):
    return np.ascontiguousarray([[float(g[0]), float(g[1]), float(g[2])] for g in grid],
                                dtype=np.float64).ravel()


def flowmap_composition(flowmaps, grid, nT, *, device_out=False):
    """Composed flow map (nx, ny, 2) from the nT intermediate flow maps (nT, nx, ny, 2): bilinear
    interpolation on `grid` = ((x0, x1, nx), (y0, y1, ny)), 0 outside the grid (the reference's
    eval_linear(..., xto.CONSTANT)), all passes fused into one kernel."""
    fa = _lib.arg_in(flowmaps)
    g = _grid6(grid)
    nx, ny = int(grid[0][2]), int(grid[1][2])
    if tuple(int(v) for v in fa.obj.shape) != (int(nT), nx, ny, 2):
        raise ValueError(f"flowmaps must have shape {(int(nT), nx, ny, 2)}")
    dev = bool(device_out or fa.on_device)
    out = _lib.alloc_out((nx, ny, 2), np.float64, dev)
    _lib.check(_lib.load().b200cs_flowmap_composition(fa.ptr, C.c_void_p(g.ctypes.data), int(nT),
                                                      out.ptr, _lib.current_stream(dev)))
    return out.obj


def _flowmap_into(dst, funcptr, t0, h, x, y, params, kwargs):
    """flowmap_grid_2D written straight into dst (a (nx, ny, 2) slice of the flowmaps array)."""
    kw = dict(kwargs)
    xa, ya, pa, ma = _lib.arg_in(x), _lib.arg_in(y), _lib.arg_in(params), _lib.mask_in(kw.pop("mask", None))
    method, rtol, atol = kw.pop("method", "dop853"), kw.pop("rtol", 1e-6), kw.pop("atol", 1e-8)
    if kw:
        raise TypeError(f"unexpected keyword arguments {sorted(kw)}")
    on_dev = _lib._is_torch(dst)
    ptr = C.c_void_p(dst.data_ptr() if on_dev else dst.ctypes.data)
    _lib.check(_lib.load().b200cs_flowmap_grid_2d(
        int(funcptr), float(t0), float(h), xa.ptr, int(xa.obj.shape[0]), ya.ptr, int(ya.obj.shape[0]),
        pa.ptr, int(pa.obj.shape[0]), _method(method), float(rtol), float(atol), ma.ptr, 0, ptr, None,
        None, None, None, _lib.current_stream(on_dev)))


def _new_flowmaps(nT, nx, ny, device):
    if device:
        import torch
        return torch.zeros((nT, nx, ny, 2), dtype=torch.float64, device="cuda")
    return np.zeros((nT, nx, ny, 2), np.float64)


def flowmap_composition_initial(funcptr, t0, T, h, x, y, grid, params, *, device_out=False, **kwargs):
    """First step of the composition method -> (flowmap0 (nx, ny, 2), flowmaps (nT, nx, ny, 2), nT)
    with nT = |round(T / h)|; flowmaps[k] is the flow map over [t0 + k h, t0 + (k + 1) h]."""
    nT = abs(round(T / h))
    nx, ny = int(grid[0][2]), int(grid[1][2])
    dev = bool(device_out or _lib._is_torch(x) and x.is_cuda)
    flowmaps = _new_flowmaps(nT, nx, ny, dev)
    t0s = np.empty(nT, np.float64)
    for k in range(nT):            # the reference's running sum t0 += h (integration.py:686-688)
        t0s[k] = t0
        t0 += h
    if nT:
        flowmap_grid_2D_series(funcptr, t0s, h, x, y, params, out=flowmaps, **kwargs)
    return flowmap_composition(flowmaps, grid, nT, device_out=dev), flowmaps, nT


def flowmap_composition_step(flowmaps, funcptr, t0, h, nT, x, y, grid, params, **kwargs):
    """Next step: drops flowmaps[0], appends the flow map over [t0, t0 + h] (integrated on the GPU,
    in place) and composes -> (flowmap_k, flowmaps).  Like the reference, `flowmaps` is updated in
    place when it is a float64 numpy array or CUDA tensor."""
    if _lib._is_torch(flowmaps):
        flowmaps[:-1] = flowmaps[1:].clone()
    else:
        flowmaps = np.asarray(flowmaps)
        if flowmaps.dtype != np.float64 or not flowmaps.flags.c_contiguous:
            flowmaps = np.ascontiguousarray(flowmaps, dtype=np.float64)
        flowmaps[:-1] = flowmaps[1:].copy()
    _flowmap_into(flowmaps[-1], funcptr, t0, h, x, y, params, kwargs)
    return flowmap_composition(flowmaps, grid, nT), flowmaps


def flowmap_composition_series(funcptr, t0, T, h, n, x, y, grid, params, *, device_out=False, **kwargs):
    """The n composed flow maps that flowmap_composition_initial followed by n - 1 calls of
    flowmap_composition_step produce (frame k covers [t0 + k h, t0 + k h + T]) -> (n, nx, ny, 2),
    in TWO launches: every intermediate map of every frame from one time-series integration
    (nT + n - 1 maps, each over one interval h), then one sliding-window composition kernel.
    The interval start times are the running sums t0, t0 + h, (t0 + h) + h, ... of the reference's
    loop, so frame 0 equals flowmap_composition_initial's result bit for bit."""
    nT = abs(round(T / h))
    nx, ny = int(grid[0][2]), int(grid[1][2])
    if nT < 1 or n < 1:
        raise ValueError("need at least one intermediate map and one frame")
    dev = bool(device_out or _lib._is_torch(x) and x.is_cuda)
    nmaps = nT + int(n) - 1
    t0s = np.empty(nmaps, np.float64)
    t = t0
    for k in range(nT):                  # flowmap_composition_initial: t0 += h per map
        t0s[k] = t
        t += h
    for k in range(1, int(n)):           # step k integrates from t0 + T + (k - 1) h in the examples
        t0s[nT + k - 1] = t0 + T + (k - 1) * h
    flowmaps = flowmap_grid_2D_series(funcptr, t0s, h, x, y, params, device_out=dev, **kwargs)
    g = _grid6(grid)
    out = _lib.alloc_out((int(n), nx, ny, 2), np.float64, dev)
    fa = _lib.arg_in(flowmaps)
    _lib.check(_lib.load().b200cs_flowmap_composition_series(fa.ptr, C.c_void_p(g.ctypes.data), nT, int(n),
                                                             out.ptr, _lib.current_stream(dev)))
    return out.obj
