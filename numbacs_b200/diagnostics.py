"""Diagnostics -- drop-in for the hot-path part of ``numbacs.diagnostics``.

ftle_grid_2D (diagnostics.py:21-65) and lavd_grid_2D (272-379) of the reference with the same
arguments and layouts; both accept numpy arrays or torch CUDA tensors (then nothing crosses PCIe)
and run as CUDA kernels of libb200cs.so.  flowmap_ftle_grid_2D is the fused convenience call for
the README workflow (flow map + FTLE with the flow map kept on the device).

C_tensor_2D (68-112), C_eig_aux_2D (115-197), C_eig_2D (200-244) and ftle_from_eig (247-269) feed
the ridge / LCS extraction: eigenvalues ascending, eigenvectors as columns with np.linalg.eigh's
(LAPACK's) sign conventions, bit-identical to the reference.
"""
import ctypes as C

import numpy as np

from . import _lib
from .flows import ScalarField
from .integration import _info_bufs, _fill_info, _method

__all__ = ["ftle_grid_2D", "ftle_grid_2D_series", "ftle_slab_2D", "lavd_grid_2D", "flowmap_ftle_grid_2D",
           "lavd_flowmap_grid_2D", "lavd_vort_sums", "C_tensor_2D", "C_eig_aux_2D", "C_eig_2D", "ftle_from_eig"]


def ftle_grid_2D(flowmap, T, dx, dy, mask=None, *, device_out=False):
    """FTLE field (nx, ny) from a flow map (nx, ny, 2) for integration time T."""
    fm, ma = _lib.arg_in(flowmap), _lib.mask_in(mask)
    if fm.obj.ndim != 3 or fm.obj.shape[2] != 2:
        raise ValueError("flowmap must have shape (nx, ny, 2)")
    nx, ny = int(fm.obj.shape[0]), int(fm.obj.shape[1])
    dev = bool(device_out or fm.on_device)
    out = _lib.alloc_out((nx, ny), np.float64, dev)
    _lib.check(_lib.load().b200cs_ftle_grid_2d(fm.ptr, nx, ny, float(T), float(dx), float(dy),
                                               ma.ptr, out.ptr, _lib.current_stream(dev)))
    return out.obj


def ftle_grid_2D_series(flowmaps, T, dx, dy, mask=None, *, device_out=False):
    """ftle_grid_2D on every frame of flowmaps (nt, nx, ny, 2) -> (nt, nx, ny), one launch; frame f
    is bit-identical to ftle_grid_2D(flowmaps[f], T, dx, dy, mask)."""
    fm, ma = _lib.arg_in(flowmaps), _lib.mask_in(mask)
    if fm.obj.ndim != 4 or fm.obj.shape[3] != 2:
        raise ValueError("flowmaps must have shape (nt, nx, ny, 2)")
    nt, nx, ny = (int(v) for v in fm.obj.shape[:3])
    dev = bool(device_out or fm.on_device)
    out = _lib.alloc_out((nt, nx, ny), np.float64, dev)
    _lib.check(_lib.load().b200cs_ftle_series_2d(fm.ptr, nt, nx, ny, float(T), float(dx), float(dy),
                                                 ma.ptr, out.ptr, _lib.current_stream(dev)))
    return out.obj


def ftle_slab_2D(flowmap, T, dx, dy, halo=(0, 0), mask=None, *, device_out=False):
    """FTLE of a row slab (nx, ny, 2) that carries halo[0] / halo[1] stencil-only rows at its
    first / last row (multi-GPU row blocks) -> (nx - halo[0] - halo[1], ny)."""
    fm, ma = _lib.arg_in(flowmap), _lib.mask_in(mask)
    nx, ny = int(fm.obj.shape[0]), int(fm.obj.shape[1])
    dev = bool(device_out or fm.on_device)
    out = _lib.alloc_out((nx - halo[0] - halo[1], ny), np.float64, dev)
    _lib.check(_lib.load().b200cs_ftle_slab_2d(fm.ptr, nx, ny, float(T), float(dx), float(dy), ma.ptr,
                                               int(halo[0]), int(halo[1]), out.ptr,
                                               _lib.current_stream(dev)))
    return out.obj


def flowmap_ftle_grid_2D(funcptr, t0, T, x, y, params, dx, dy, method="dop853", rtol=1e-6,
                         atol=1e-8, mask=None, *, return_flowmap=True, device_out=False,
                         info=None, halo=(0, 0)):
    """flowmap_grid_2D + ftle_grid_2D in one call.  Returns (flowmap or None, ftle).

    halo=(lo, hi): the first / last row of x is a stencil-only halo row (multi-GPU row blocks):
    it is integrated but its FTLE row is not produced."""
    xa, ya, pa, ma = _lib.arg_in(x), _lib.arg_in(y), _lib.arg_in(params), _lib.mask_in(mask)
    nx, ny = int(xa.obj.shape[0]), int(ya.obj.shape[0])
    dev = bool(device_out or xa.on_device or ya.on_device)
    fm = _lib.alloc_out((nx, ny, 2), np.float64, dev) if return_flowmap else _lib.Arg(None, None, False)
    ftle = _lib.alloc_out((nx - halo[0] - halo[1], ny), np.float64, dev)
    status, _, stats = _info_bufs(info, (nx, ny), dev)
    _lib.check(_lib.load().b200cs_flowmap_ftle_grid_2d(
        int(funcptr), float(t0), float(T), xa.ptr, nx, ya.ptr, ny, pa.ptr, int(pa.obj.shape[0]),
        _method(method), float(rtol), float(atol), ma.ptr, float(dx), float(dy), int(halo[0]),
        int(halo[1]), fm.ptr, ftle.ptr, status.ptr, stats.ptr, _lib.current_stream(dev)))
    if info is not None:
        info["status"], info["stats"] = status.obj, stats.obj
    return fm.obj, ftle.obj


def lavd_grid_2D(flowmap_n, tspan, T, vort_interp, xrav, yrav, period_x=0.0, period_y=0.0,
                 mask=None, *, device_out=False, vort_avg=None):
    """LAVD field (nx, ny) from trajectories flowmap_n (nx, ny, n, 2) at times tspan[n].

    `vort_interp` must be a ScalarField from numbacs_b200.flows.get_callable_scalar(_linear).
    `T` is unused, as in the reference.  `vort_avg` (optional, [n]) supplies precomputed spatial
    means (used by the sharded multi-GPU driver)."""
    if not isinstance(vort_interp, ScalarField):
        raise NotImplementedError(
            "vort_interp must come from numbacs_b200.flows.get_callable_scalar / "
            "get_callable_scalar_linear: an arbitrary jit-callable cannot run on the GPU")
    fm, ts, ma = _lib.arg_in(flowmap_n), _lib.arg_in(tspan), _lib.mask_in(mask)
    if fm.obj.ndim != 4 or fm.obj.shape[3] != 2:
        raise ValueError("flowmap_n must have shape (nx, ny, n, 2)")
    nx, ny, n = (int(v) for v in fm.obj.shape[:3])
    xr, yr = _lib.arg_in(xrav), _lib.arg_in(yrav)
    nrav = int(xr.obj.shape[0])
    dev = bool(device_out or fm.on_device)
    out = _lib.alloc_out((nx, ny), np.float64, dev)
    if vort_avg is not None:
        va = _lib.arg_in(vort_avg)
        va_ptr, va_in = va.ptr, 1
    else:
        va = None
        va_ptr, va_in = None, 0
    _lib.check(_lib.load().b200cs_lavd_grid_2d(
        fm.ptr, nx, ny, n, ts.ptr, vort_interp.handle, xr.ptr, yr.ptr, nrav, float(period_x),
        float(period_y), ma.ptr, va_ptr, va_in, out.ptr, _lib.current_stream(dev)))
    return out.obj


def lavd_flowmap_grid_2D(funcptr, t0, T, x, y, params, vort_interp, n=50, method="dop853", rtol=1e-6,
                         atol=1e-8, period_x=0.0, period_y=0.0, mask=None, *, vort_avg=None,
                         return_flowmap=False, device_out=False, info=None):
    """flowmap_n_grid_2D + lavd_grid_2D fused: the LAVD is accumulated along each trajectory while
    it is integrated, so the (nx, ny, n, 2) array is never stored.  Returns (lavd, tspan) or
    (lavd, tspan, final_flowmap).  Same result as the two reference calls
    (integration.py:467-533, diagnostics.py:272-379) up to the order of the Simpson summation.
    The vorticity is evaluated on slabs contracted over time at the n output times (16 taps per
    evaluation instead of 64; include/b200cs.h), the spatial means through per-axis weight sums."""
    if not isinstance(vort_interp, ScalarField):
        raise NotImplementedError("vort_interp must come from numbacs_b200.flows.get_callable_scalar(_linear)")
    xa, ya, pa, ma = _lib.arg_in(x), _lib.arg_in(y), _lib.arg_in(params), _lib.mask_in(mask)
    nx, ny = int(xa.obj.shape[0]), int(ya.obj.shape[0])
    dev = bool(device_out or xa.on_device or ya.on_device)
    lavd = _lib.alloc_out((nx, ny), np.float64, dev)
    fm = _lib.alloc_out((nx, ny, 2), np.float64, dev) if return_flowmap else _lib.Arg(None, None, False)
    tspan = np.empty(int(n), np.float64)
    status, _, stats = _info_bufs(info, (nx, ny), dev)
    va = _lib.arg_in(vort_avg) if vort_avg is not None else _lib.Arg(None, None, False)
    _lib.check(_lib.load().b200cs_lavd_flowmap_grid_2d(
        int(funcptr), float(t0), float(T), xa.ptr, nx, ya.ptr, ny, pa.ptr, int(pa.obj.shape[0]),
        _method(method), float(rtol), float(atol), ma.ptr, int(n), vort_interp.handle, float(period_x),
        float(period_y), va.ptr, lavd.ptr, fm.ptr, C.c_void_p(tspan.ctypes.data), status.ptr, stats.ptr,
        _lib.current_stream(dev)))
    if info is not None:
        info["status"], info["stats"] = status.obj, stats.obj
    return (lavd.obj, tspan, fm.obj) if return_flowmap else (lavd.obj, tspan)


def lavd_vort_sums(vort_interp, tspan, xrav, yrav, *, device_out=False):
    """sums[k] = sum_q vort(tspan[k], xrav[q], yrav[q]): the un-normalised spatial mean of
    lavd_grid_2D (diagnostics.py:324-331).  The multi-GPU driver all-reduces these n doubles."""
    if not isinstance(vort_interp, ScalarField):
        raise NotImplementedError("vort_interp must come from numbacs_b200.flows.get_callable_scalar(_linear)")
    ts, xr, yr = _lib.arg_in(tspan), _lib.arg_in(xrav), _lib.arg_in(yrav)
    n, nrav = int(ts.obj.shape[0]), int(xr.obj.shape[0])
    dev = bool(device_out or xr.on_device)
    out = _lib.alloc_out((n,), np.float64, dev)
    _lib.check(_lib.load().b200cs_lavd_vort_sums(vort_interp.handle, ts.ptr, n, xr.ptr, yr.ptr, nrav,
                                                 out.ptr, _lib.current_stream(dev)))
    return out.obj


def _aux_in(flowmap_aux):
    fa = _lib.arg_in(flowmap_aux)
    if fa.obj.ndim != 4 or fa.obj.shape[3] != 2 or fa.obj.shape[2] not in (4, 5):
        raise ValueError("flowmap_aux must have shape (nx, ny, 4 or 5, 2)")
    return fa, int(fa.obj.shape[0]), int(fa.obj.shape[1]), int(fa.obj.shape[2])


def C_tensor_2D(flowmap_aux, dx, dy, h=1e-5, mask=None, *, device_out=False):
    """(C11, C12, C22) of the Cauchy-Green tensor from the aux-grid flow map -> (nx, ny, 3);
    zero outside [2, nx-2) x [2, ny-2) and where masked.  dx, dy are unused, as in the reference."""
    fa, nx, ny, n_aux = _aux_in(flowmap_aux)
    ma = _lib.mask_in(mask)
    dev = bool(device_out or fa.on_device)
    out = _lib.alloc_out((nx, ny, 3), np.float64, dev)
    _lib.check(_lib.load().b200cs_c_tensor_2d(fa.ptr, nx, ny, n_aux, float(dx), float(dy), float(h),
                                              ma.ptr, out.ptr, _lib.current_stream(dev)))
    return out.obj


def C_eig_aux_2D(flowmap_aux, dx, dy, h=1e-5, eig_main=True, mask=None, *, device_out=False):
    """Eigenvalues (nx, ny, 2) and eigenvectors (nx, ny, 2, 2) of the Cauchy-Green tensor from the
    aux-grid flow map; with eig_main the eigenvalues come from the main-grid stencil."""
    fa, nx, ny, n_aux = _aux_in(flowmap_aux)
    if eig_main and n_aux != 5:
        raise ValueError("eig_main=True needs flowmap_aux with the centre point (n_aux = 5)")
    ma = _lib.mask_in(mask)
    dev = bool(device_out or fa.on_device)
    vals = _lib.alloc_out((nx, ny, 2), np.float64, dev)
    vecs = _lib.alloc_out((nx, ny, 2, 2), np.float64, dev)
    _lib.check(_lib.load().b200cs_c_eig_aux_2d(fa.ptr, nx, ny, n_aux, float(dx), float(dy), float(h),
                                               int(bool(eig_main)), ma.ptr, vals.ptr, vecs.ptr,
                                               _lib.current_stream(dev)))
    return vals.obj, vecs.obj


def C_eig_2D(flowmap, dx, dy, mask=None, *, device_out=False, ftle_T=None):
    """Eigenvalues (nx, ny, 2) and eigenvectors (nx, ny, 2, 2) of the Cauchy-Green tensor from a
    flow map (nx, ny, 2).  With ftle_T = T the FTLE field ftle_from_eig(eigvals[:, :, 1], T) is
    produced in the same pass and returned as a third array (bit-identical to the separate call)."""
    fm, ma = _lib.arg_in(flowmap), _lib.mask_in(mask)
    if fm.obj.ndim != 3 or fm.obj.shape[2] != 2:
        raise ValueError("flowmap must have shape (nx, ny, 2)")
    nx, ny = int(fm.obj.shape[0]), int(fm.obj.shape[1])
    dev = bool(device_out or fm.on_device)
    vals = _lib.alloc_out((nx, ny, 2), np.float64, dev)
    vecs = _lib.alloc_out((nx, ny, 2, 2), np.float64, dev)
    if ftle_T is not None:
        ft = _lib.alloc_out((nx, ny), np.float64, dev)
        _lib.check(_lib.load().b200cs_c_eig_ftle_2d(fm.ptr, nx, ny, float(dx), float(dy), float(ftle_T), ma.ptr,
                                                    vals.ptr, vecs.ptr, ft.ptr, _lib.current_stream(dev)))
        return vals.obj, vecs.obj, ft.obj
    _lib.check(_lib.load().b200cs_c_eig_2d(fm.ptr, nx, ny, float(dx), float(dy), ma.ptr, vals.ptr,
                                           vecs.ptr, _lib.current_stream(dev)))
    return vals.obj, vecs.obj


def ftle_from_eig(eigval_max, T, *, device_out=False):
    """FTLE from the largest Cauchy-Green eigenvalue: log(eigval_max) / (2|T|) where > 1, else 0.
    A last-axis slice such as eigvals[:, :, 1] is read in place (strided)."""
    e, stride = _lib.strided_in(eigval_max)
    shape = tuple(int(v) for v in eigval_max.shape)
    dev = bool(device_out or e.on_device)
    out = _lib.alloc_out(shape, np.float64, dev)
    n = int(np.prod(shape)) if shape else 1
    _lib.check(_lib.load().b200cs_ftle_from_eig(e.ptr, n, stride, float(T), out.ptr,
                                                _lib.current_stream(dev)))
    return out.obj
