"""numbalsoda shim (test / baseline infrastructure): the names the reference imports.

    from numbalsoda import lsoda_sig, dop853, lsoda          (flows.py:4, integration.py:2)

dop853(funcptr, u0, t_eval, data=..., rtol=..., atol=...) -> (usol[nt, neq], success) is callable
from numba nopython code (the reference calls it inside @njit(parallel=True) prange loops,
integration.py:49, 108, 169, 520) and runs Hairer's DOP853 in native code with the RHS given as a C
function pointer: oracle_dop853_callback of oracle/libnumbacs_oracle.so.
"""
import ctypes as _ct
import os as _os
import subprocess as _sp

import numpy as np
from numba import njit, types

_HERE = _os.path.dirname(_os.path.abspath(__file__))
_ORACLE = _os.path.normpath(_os.path.join(_HERE, "..", ".."))
_SO = _os.path.join(_ORACLE, "libnumbacs_oracle.so")
if not _os.path.exists(_SO):
    _sp.check_call(["make", "-s", "-C", _ORACLE])
_lib = _ct.CDLL(_SO)

# void rhs(double t, double *y, double *dy, double *p)
lsoda_sig = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                       types.CPointer(types.double))

_dop853_cb = _lib.oracle_dop853_callback
_dop853_cb.restype = _ct.c_int
_dop853_cb.argtypes = [_ct.c_void_p, _ct.c_int, _ct.c_void_p, _ct.c_void_p, _ct.c_int64, _ct.c_double,
                       _ct.c_double, _ct.c_void_p, _ct.c_void_p]


@njit
def dop853(funcptr, u0, t_eval, data=np.array([0.0], np.float64), rtol=1.0e-3, atol=1.0e-6):
    neq = u0.shape[0]
    nt = t_eval.shape[0]
    u0c = np.ascontiguousarray(u0)
    tc = np.ascontiguousarray(t_eval)
    dc = np.ascontiguousarray(data)
    usol = np.empty((nt, neq), np.float64)
    rc = _dop853_cb(funcptr, neq, u0c.ctypes.data, tc.ctypes.data, nt, rtol, atol, dc.ctypes.data,
                    usol.ctypes.data)
    return usol, rc == 1


@njit
def lsoda(funcptr, u0, t_eval, data=np.array([0.0], np.float64), rtol=1.0e-3, atol=1.0e-6):
    # the reference compiles both branches of `method`; LSODA itself is out of scope
    usol = np.full((t_eval.shape[0], u0.shape[0]), np.nan)
    return usol, False
