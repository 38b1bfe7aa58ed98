"""ctypes binding of libb200cs.so (the C-ABI declared in include/b200cs.h).

This is the whole Python<->CUDA boundary: plain pointers and sizes.  numpy arrays are passed as
host pointers (the library stages them), torch CUDA tensors as device pointers.  There is no CPU
fallback: if the shared library has not been built, or no GPU is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200CS_LIB selects an experimental build of the same library (tools/build_variant.py, A/B kernel
# measurements); the product path is the in-tree libb200cs.so
LIB_PATH = os.environ.get("B200CS_LIB") or os.path.join(_HERE, "libb200cs.so")

E_INVALID, E_CUDA, E_HANDLE, E_UNSUPPORTED = -1, -2, -3, -4
FLOW_KINDS = {"double_gyre": 0, "bickley_jet": 1, "abc": 2}
EXTRAP = {"constant": 0, "linear": 1, "nearest": 2}
METHOD_DOP853 = 0

_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_ip = C.POINTER(C.c_int)

# name -> argtypes, exactly the prototypes of include/b200cs.h (restype is int unless noted)
PROTOTYPES = {
    "b200cs_version": [],
    "b200cs_device_count": [_ip],
    "b200cs_flow_create_analytic": [_i, _ip],
    "b200cs_flow_create_spline": [_vp, _vp, _vp, _i, _i, _d, _ip],
    "b200cs_flow_create_linear": [_vp, _vp, _vp, _i, _i, _d, _ip],
    "b200cs_scalar_create": [_vp, _vp, _i, _i, _ip],
    "b200cs_flow_destroy": [_i],
    "b200cs_flow_info": [_i, _ip, _ip, _ip],
    "b200cs_flow_out_of_grid": [_i, _vp, _i, _vp],
    "b200cs_prefilter_3d": [_vp, _i64, _i64, _i64, _vp, _vp],
    "b200cs_scalar_eval": [_i, _vp, _i64, _vp, _vp],
    "b200cs_velocity_eval": [_i, _vp, _i64, _vp, _vp],
    "b200cs_curl_func_tspan": [_i, _vp, _i64, _vp, _i64, _vp, _i64, _d, _vp, _vp],
    "b200cs_flow_rhs": [_i, _vp, _vp, _i64, _vp, _i, _vp, _vp],
    "b200cs_flowmap_grid_2d": [_i, _d, _d, _vp, _i64, _vp, _i64, _vp, _i, _i, _d, _d, _vp, _i,
                               _vp, _vp, _vp, _vp, _vp, _vp],
    "b200cs_flowmap_pts": [_i, _d, _d, _vp, _i64, _i, _vp, _i, _i, _d, _d, _vp, _i, _vp, _vp,
                           _vp, _vp, _vp, _vp],
    "b200cs_ftle_grid_2d": [_vp, _i64, _i64, _d, _d, _d, _vp, _vp, _vp],
    "b200cs_ftle_slab_2d": [_vp, _i64, _i64, _d, _d, _d, _vp, _i, _i, _vp, _vp],
    "b200cs_flowmap_ftle_grid_2d": [_i, _d, _d, _vp, _i64, _vp, _i64, _vp, _i, _i, _d, _d, _vp,
                                    _d, _d, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "b200cs_lavd_grid_2d": [_vp, _i64, _i64, _i64, _vp, _i, _vp, _vp, _i64, _d, _d, _vp, _vp, _i,
                            _vp, _vp],
    "b200cs_lavd_flowmap_grid_2d": [_i, _d, _d, _vp, _i64, _vp, _i64, _vp, _i, _i, _d, _d, _vp, _i, _i, _d, _d,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "b200cs_lavd_vort_sums": [_i, _vp, _i64, _vp, _vp, _i64, _vp, _vp],
    "b200cs_flowmap_aux_grid_2d": [_i, _d, _d, _vp, _i64, _vp, _i64, _vp, _i, _d, _i, _i, _i, _d, _d,
                                   _vp, _vp, _vp, _vp, _vp, _vp],
    "b200cs_flowmap_grid_2d_series": [_i, _vp, _i64, _d, _vp, _i64, _vp, _i64, _vp, _i, _i, _d, _d, _vp, _vp,
                                      _vp, _vp, _vp, _vp],
    "b200cs_ftle_series_2d": [_vp, _i64, _i64, _i64, _d, _d, _d, _vp, _vp, _vp],
    "b200cs_c_tensor_2d": [_vp, _i64, _i64, _i, _d, _d, _d, _vp, _vp, _vp],
    "b200cs_c_eig_2d": [_vp, _i64, _i64, _d, _d, _vp, _vp, _vp, _vp],
    "b200cs_c_eig_ftle_2d": [_vp, _i64, _i64, _d, _d, _d, _vp, _vp, _vp, _vp, _vp],
    "b200cs_c_eig_aux_2d": [_vp, _i64, _i64, _i, _d, _d, _d, _i, _vp, _vp, _vp, _vp],
    "b200cs_ftle_from_eig": [_vp, _i64, _i64, _d, _vp, _vp],
    "b200cs_ftle_ridge_pts": [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _d, _d, _d, _d, _vp, _vp, _vp,
                              _vp, _i64, _vp, _vp],
    "b200cs_ftle_ridges": [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _d, _d, _d, _d, _vp, _vp, _i64, _vp,
                           _vp],
    "b200cs_link_ridge_pts": [_vp, _vp, _vp, _i64, _i64, _d, _d, _d, _vp, _i64, _vp, _vp, _vp, _i64, _vp],
    "b200cs_order_ridges": [_vp, _i64, _vp, _vp, _vp, _i64, _d, _d, _i64, _vp, _vp, _vp],
    "b200cs_flowmap_composition": [_vp, _vp, _i64, _vp, _vp],
    "b200cs_binary_mask_dilation": [_vp, _i64, _i64, _i, _vp, _vp],
    "b200cs_flowmap_composition_series": [_vp, _vp, _i64, _i64, _vp, _vp],
    "b200cs_order_stats": [_vp, _i64, _i64, _vp, _vp],
    "b200cs_fp64_peak": [_i, C.POINTER(_d), C.POINTER(_d)],
}

_lib = None
_lib_path = None
# the parity-calibration build of the same sources (B200CS_STRICT, csrc/dop853.cuh): test infrastructure
STRICT_LIB_PATH = os.path.join(_HERE, "libb200cs_strict.so")


def _open(path):
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build the CUDA library first "
            "(python -m numbacs_b200._build, or __graft_entry__.build()). "
            "numbacs_b200 has no CPU fallback.")
    L = C.CDLL(path)
    L.b200cs_last_error.restype = C.c_char_p
    L.b200cs_last_error.argtypes = []
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    return L


def load():
    """Load libb200cs.so; raises if it has not been built (python -m numbacs_b200._build)."""
    global _lib, _lib_path
    if _lib is None:
        _lib = _open(LIB_PATH)
        _lib_path = LIB_PATH
    return _lib


def library_info():
    """Which shared library the API is bound to, and whether B200CS_LIB replaced the product one."""
    load()
    return {"path": _lib_path, "overridden_by_B200CS_LIB": bool(os.environ.get("B200CS_LIB")),
            "version": int(_lib.b200cs_version())}


class use_library:
    """Context manager for tests / A-B tools: bind the whole Python API to another build of the
    library (e.g. STRICT_LIB_PATH).  Flow handles belong to the library that created them, so
    create the flows inside the block."""

    def __init__(self, path):
        self.path = path

    def __enter__(self):
        global _lib, _lib_path
        self.saved = (_lib, _lib_path)
        _lib, _lib_path = _open(self.path), self.path
        return _lib

    def __exit__(self, *exc):
        global _lib, _lib_path
        _lib, _lib_path = self.saved
        return False


def check(rc):
    if rc == 0:
        return
    msg = load().b200cs_last_error().decode("utf-8", "replace")
    if rc == E_HANDLE:
        raise NotImplementedError(
            msg + " -- numbacs_b200 integrates only flows created by its own get_predefined_flow / "
            "get_flow_2D (device implementations); user-written numba cfuncs cannot run on the "
            "GPU and there is no CPU fallback.")
    if rc == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == E_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def _is_torch(a):
    return type(a).__module__.split(".")[0] == "torch"


class Arg:
    """A marshalled array argument: .ptr for ctypes, keeps the backing object alive."""
    __slots__ = ("obj", "ptr", "on_device")

    def __init__(self, obj, ptr, on_device):
        self.obj, self.ptr, self.on_device = obj, ptr, on_device


def arg_in(a, dtype=np.float64):
    """numpy / array-like -> host pointer; torch tensor -> its (host or device) pointer."""
    if a is None:
        return Arg(None, None, False)
    if _is_torch(a):
        import torch
        tdt = {np.float64: torch.float64, np.uint8: torch.uint8, np.bool_: torch.bool}[dtype]
        t = a
        if dtype is np.uint8 and t.dtype == torch.bool:
            t = t.view(torch.uint8) if t.is_contiguous() else t.contiguous().view(torch.uint8)
        elif t.dtype != tdt:
            t = t.to(tdt)
        t = t.contiguous()
        return Arg(t, C.c_void_p(t.data_ptr()), t.is_cuda)
    arr = np.ascontiguousarray(a, dtype=dtype)
    return Arg(arr, C.c_void_p(arr.ctypes.data), False)


def _elem_strides(a):
    """Strides in elements for numpy arrays and torch tensors alike."""
    if _is_torch(a):
        return tuple(int(v) for v in a.stride())
    return tuple(int(v) // a.itemsize for v in a.strides)


def strided_in(a):
    """An array whose elements lie at a constant element stride in memory (a contiguous array, or
    a last-axis slice like eigvals[:, :, 1] of a contiguous one) -> (Arg, stride); anything else
    is copied to a contiguous buffer first."""
    is_t = _is_torch(a)
    if not is_t:
        a = np.asarray(a)
    ok = (a.dtype == np.float64) if not is_t else (str(a.dtype) == "torch.float64")
    if ok and a.ndim >= 1 and all(int(v) > 0 for v in a.shape):
        st, shape = _elem_strides(a), tuple(int(v) for v in a.shape)
        s = st[-1]
        if s >= 1 and all(st[d] == st[d + 1] * shape[d + 1] for d in range(a.ndim - 1)):
            ptr = a.data_ptr() if is_t else a.ctypes.data
            return Arg(a, C.c_void_p(ptr), bool(is_t and a.is_cuda)), s
    return arg_in(a), 1


def mask_in(mask):
    if mask is None:
        return Arg(None, None, False)
    if _is_torch(mask):
        return arg_in(mask, np.uint8)
    m = np.ascontiguousarray(mask)
    if m.dtype != np.bool_ and m.dtype != np.uint8:
        m = m.astype(np.bool_)
    m = m.view(np.uint8)
    return Arg(m, C.c_void_p(m.ctypes.data), False)


def alloc_out(shape, dtype, device):
    """Output buffer: numpy (host) or torch CUDA tensor (device)."""
    if device:
        import torch
        tdt = {np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64,
               np.bool_: torch.bool}[dtype]
        t = torch.empty(shape, dtype=tdt, device="cuda")
        return Arg(t, C.c_void_p(t.data_ptr()), True)
    arr = np.empty(shape, dtype=dtype)
    return Arg(arr, C.c_void_p(arr.ctypes.data), False)


def current_stream(device):
    """torch's current CUDA stream when tensors are involved, else the default stream."""
    if device:
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return None


def device_count():
    n = C.c_int(0)
    check(load().b200cs_device_count(C.byref(n)))
    return n.value


def fp64_peak(iters=20000):
    """Measured FP64 FMA peak of the current GPU in TFLOP/s (register-resident DFMA chains)."""
    tf, ms = C.c_double(0.0), C.c_double(0.0)
    check(load().b200cs_fp64_peak(int(iters), C.byref(tf), C.byref(ms)))
    return tf.value, ms.value
