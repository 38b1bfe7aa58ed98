"""Mask helpers -- drop-in for the mask pipeline of ``numbacs.utils`` that feeds the ``mask``
argument of the flow-map / FTLE / Cauchy-Green kernels.

binary_mask_dilation (utils.py:1923-1985) runs as a CUDA kernel (the reference recommends dilating
the mask before C_eig_2D / C_eig_aux_2D so that stencils never straddle masked data);
fill_nans_and_get_mask (utils.py:1819-1855) is a one-off in-place data preparation step on the
caller's arrays (numpy or torch), exactly the reference's assignments.
"""
import numpy as np

from . import _lib

__all__ = ["binary_mask_dilation", "fill_nans_and_get_mask", "curl_func_tspan"]


def curl_func_tspan(fnc, t, x, y, h=1e-3, *, device_out=False):
    """Curl (vorticity) of the velocity callable `fnc` over times t and the grid (x, y) by central
    differences of spacing h -> (nt, nx, ny)   (utils.py:570-608; the vorticity field of
    examples/elliptic_lcs/plot_qge_elliptic_lcs.py:60).  `fnc` must come from
    numbacs_b200.flows.get_callable_2D: an arbitrary jit-callable cannot run on the GPU."""
    from .flows import VelocityField
    if not isinstance(fnc, VelocityField):
        raise NotImplementedError("fnc must come from numbacs_b200.flows.get_callable_2D: an arbitrary "
                                  "jit-callable cannot run on the GPU and there is no CPU fallback")
    ta, xa, ya = _lib.arg_in(t), _lib.arg_in(x), _lib.arg_in(y)
    nt, nx, ny = int(ta.obj.shape[0]), int(xa.obj.shape[0]), int(ya.obj.shape[0])
    dev = bool(device_out or ta.on_device or xa.on_device or ya.on_device)
    out = _lib.alloc_out((nt, nx, ny), np.float64, dev)
    _lib.check(_lib.load().b200cs_curl_func_tspan(fnc.handle, ta.ptr, nt, xa.ptr, nx, ya.ptr, ny, float(h),
                                                  out.ptr, _lib.current_stream(dev)))
    return out.obj


def binary_mask_dilation(mask, corners=False, *, device_out=False):
    """Binary dilation of a (nx, ny) boolean mask with the 4 cardinal neighbours (corners=True:
    plus the 4 diagonal ones)."""
    ma = _lib.mask_in(mask)
    if ma.obj is None or ma.obj.ndim != 2:
        raise ValueError("mask must have shape (nx, ny)")
    nx, ny = int(ma.obj.shape[0]), int(ma.obj.shape[1])
    dev = bool(device_out or ma.on_device)
    out = _lib.alloc_out((nx, ny), np.bool_, dev)
    _lib.check(_lib.load().b200cs_binary_mask_dilation(ma.ptr, nx, ny, int(bool(corners)), out.ptr,
                                                       _lib.current_stream(dev)))
    return out.obj


def fill_nans_and_get_mask(arrs, fill_value=0.0):
    """Boolean (nx, ny) mask of the NaNs of arrs[0][0]; every array (nt, nx, ny) of `arrs` is set
    to 0 there IN PLACE for all times (the reference ignores `fill_value` too, utils.py:1846-1852).
    Returns (*arrs, mask)."""
    arr0 = arrs[0]
    if _lib._is_torch(arr0):
        import torch
        mask = torch.isnan(arr0[0])
    else:
        mask = np.isnan(arr0[0])
    out = []
    for arr in arrs:
        arr[:, mask] = 0.0
        out.append(arr)
    return (*out, mask)
