"""interpolation.splines shim (test / baseline infrastructure): the names the reference imports.

    from interpolation.splines import UCGrid, prefilter, eval_spline, eval_linear
    from interpolation.splines import extrap_options as xto      (flows.py:5-6, integration.py:4)

Uniform cartesian grids, natural-boundary cubic B-spline prefilter, tensor-product cubic / linear
evaluation with the three extrapolation modes, in 2 and 3 dimensions, for one point (-> scalar) or
an (N, d) array of points (-> (N,) array).  eval_spline / eval_linear are numba @overload's, so they
compile inside the reference's @cfunc right-hand sides and @njit callables.

Interior behaviour is pinned by the reference's tables (tests/test_flows.py:74-133, 218-229).  The
extrapolation modes follow the package's documented meaning (nearest = clamp the coordinate,
constant = 0 outside the grid, linear = continue the blending weights linearly); no reference test
leaves the grid, so they are unpinned.
"""
import numpy as np
from numba import njit, types
from numba.extending import overload


def UCGrid(*axes):
    """((a0, b0, n0), (a1, b1, n1), ...) with float bounds and int sizes."""
    return tuple((float(a), float(b), int(n)) for a, b, n in axes)


class _ExtrapOptions:
    CONSTANT = 0
    LINEAR = 1
    NEAREST = 2


extrap_options = _ExtrapOptions()


# ------------------------------------------------------------------ prefilter

def _prefilter_axis(d, axis):
    """natural-BC cubic B-spline coefficients along one axis: n data -> n + 2 coefficients with
    c[0] - 2 c[1] + c[2] = 0, (c[i] + 4 c[i+1] + c[i+2]) / 6 = d[i], c[n-1] - 2 c[n] + c[n+1] = 0."""
    from scipy.linalg import solve_banded
    d = np.moveaxis(d, axis, 0)
    n = d.shape[0]
    rhs = np.zeros((n + 2,) + d.shape[1:])
    rhs[1:n + 1] = d
    A = np.zeros((n + 2, n + 2))
    A[0, :3] = (1.0, -2.0, 1.0)
    for i in range(n):
        A[i + 1, i:i + 3] = (1.0 / 6.0, 4.0 / 6.0, 1.0 / 6.0)
    A[n + 1, n - 1:n + 2] = (1.0, -2.0, 1.0)
    c = np.linalg.solve(A, rhs.reshape(n + 2, -1)).reshape(rhs.shape)
    return np.moveaxis(c, 0, axis)


def prefilter(grid, V, out=None, k=3):
    if k != 3:
        raise NotImplementedError("the shim implements cubic splines (k = 3) only")
    c = np.asarray(V, dtype=np.float64)
    for axis in range(c.ndim):
        c = _prefilter_axis(c, axis)
    c = np.ascontiguousarray(c)
    if out is not None:
        out[...] = c
        return out
    return c


# ------------------------------------------------------------------ evaluation kernels

@njit
def _locate(a, b, n, x):
    delta = (b - a) / (n - 1)
    d = x - a
    fi = np.floor(d / delta)
    i = 0 if fi < 0.0 else (n - 2 if fi > n - 2 else int(fi))
    return i, (d - i * delta) / delta


@njit
def _cubic_weights(lam, linear_ext):
    P = np.empty(4)
    if linear_ext and lam < 0.0:
        P[0] = -0.5 * lam + 1.0 / 6.0
        P[1] = 4.0 / 6.0
        P[2] = 0.5 * lam + 1.0 / 6.0
        P[3] = 0.0
    elif linear_ext and lam > 1.0:
        m = lam - 1.0
        P[0] = 0.0
        P[1] = -0.5 * m + 1.0 / 6.0
        P[2] = 4.0 / 6.0
        P[3] = 0.5 * m + 1.0 / 6.0
    else:
        l2 = lam * lam
        l3 = l2 * lam
        P[0] = (-1.0 / 6.0) * l3 + 0.5 * l2 - 0.5 * lam + 1.0 / 6.0
        P[1] = 0.5 * l3 - l2 + 4.0 / 6.0
        P[2] = -0.5 * l3 + 0.5 * l2 + 0.5 * lam + 1.0 / 6.0
        P[3] = (1.0 / 6.0) * l3
    return P


@njit
def _coord(a, b, x, mode):
    """-> (inside, x'): mode 0 = constant (outside -> value 0), 2 = nearest (clamp)."""
    if mode == 0:
        if x < a or x > b:
            return False, x
    elif mode == 2:
        x = min(max(x, a), b)
    return True, x


@njit
def _spline3(grid, C, p0, p1, p2, mode):
    ok0, x0 = _coord(grid[0][0], grid[0][1], p0, mode)
    ok1, x1 = _coord(grid[1][0], grid[1][1], p1, mode)
    ok2, x2 = _coord(grid[2][0], grid[2][1], p2, mode)
    if not (ok0 and ok1 and ok2):
        return 0.0
    i0, l0 = _locate(grid[0][0], grid[0][1], grid[0][2], x0)
    i1, l1 = _locate(grid[1][0], grid[1][1], grid[1][2], x1)
    i2, l2 = _locate(grid[2][0], grid[2][1], grid[2][2], x2)
    lin = mode == 1
    P0, P1, P2 = _cubic_weights(l0, lin), _cubic_weights(l1, lin), _cubic_weights(l2, lin)
    acc0 = 0.0
    for a in range(4):
        acc1 = 0.0
        for b in range(4):
            acc2 = 0.0
            for c in range(4):
                acc2 += P2[c] * C[i0 + a, i1 + b, i2 + c]
            acc1 += P1[b] * acc2
        acc0 += P0[a] * acc1
    return acc0


@njit
def _spline2(grid, C, p0, p1, mode):
    ok0, x0 = _coord(grid[0][0], grid[0][1], p0, mode)
    ok1, x1 = _coord(grid[1][0], grid[1][1], p1, mode)
    if not (ok0 and ok1):
        return 0.0
    i0, l0 = _locate(grid[0][0], grid[0][1], grid[0][2], x0)
    i1, l1 = _locate(grid[1][0], grid[1][1], grid[1][2], x1)
    lin = mode == 1
    P0, P1 = _cubic_weights(l0, lin), _cubic_weights(l1, lin)
    acc0 = 0.0
    for a in range(4):
        acc1 = 0.0
        for b in range(4):
            acc1 += P1[b] * C[i0 + a, i1 + b]
        acc0 += P0[a] * acc1
    return acc0


@njit
def _linear3(grid, F, p0, p1, p2, mode):
    ok0, x0 = _coord(grid[0][0], grid[0][1], p0, mode)
    ok1, x1 = _coord(grid[1][0], grid[1][1], p1, mode)
    ok2, x2 = _coord(grid[2][0], grid[2][1], p2, mode)
    if not (ok0 and ok1 and ok2):
        return 0.0
    i0, l0 = _locate(grid[0][0], grid[0][1], grid[0][2], x0)
    i1, l1 = _locate(grid[1][0], grid[1][1], grid[1][2], x1)
    i2, l2 = _locate(grid[2][0], grid[2][1], grid[2][2], x2)
    v = 0.0
    for a in range(2):
        wa = l0 if a else 1.0 - l0
        va = 0.0
        for b in range(2):
            wb = l1 if b else 1.0 - l1
            va += wb * ((1.0 - l2) * F[i0 + a, i1 + b, i2] + l2 * F[i0 + a, i1 + b, i2 + 1])
        v += wa * va
    return v


@njit
def _linear2(grid, F, p0, p1, mode):
    ok0, x0 = _coord(grid[0][0], grid[0][1], p0, mode)
    ok1, x1 = _coord(grid[1][0], grid[1][1], p1, mode)
    if not (ok0 and ok1):
        return 0.0
    i0, l0 = _locate(grid[0][0], grid[0][1], grid[0][2], x0)
    i1, l1 = _locate(grid[1][0], grid[1][1], grid[1][2], x1)
    v = 0.0
    for a in range(2):
        wa = l0 if a else 1.0 - l0
        v += wa * ((1.0 - l1) * F[i0 + a, i1] + l1 * F[i0 + a, i1 + 1])
    return v


@njit
def _mode_of(extrap_mode):
    if extrap_mode == "constant":
        return 0
    if extrap_mode == "linear":
        return 1
    return 2


# ------------------------------------------------------------------ public, numba-overloaded

def eval_spline(grid, C, points, out=None, k=3, diff="None", extrap_mode="linear"):
    """Interpreter entry (the reference also calls these outside numba, integration.py:633-640);
    compiled code goes through the @overload below."""
    return _py_eval_spline(grid, np.ascontiguousarray(C), np.asarray(points, dtype=np.float64), extrap_mode)


def eval_linear(grid, V, points, extrap_mode=1):
    return _py_eval_linear(grid, np.ascontiguousarray(V), np.asarray(points, dtype=np.float64), extrap_mode)


@overload(eval_spline)
def _ov_eval_spline(grid, C, points, out=None, k=3, diff="None", extrap_mode="linear"):
    nd = len(grid)
    if points.ndim == 1:
        if nd == 3:
            def impl(grid, C, points, out=None, k=3, diff="None", extrap_mode="linear"):
                return _spline3(grid, C, points[0], points[1], points[2], _mode_of(extrap_mode))
        else:
            def impl(grid, C, points, out=None, k=3, diff="None", extrap_mode="linear"):
                return _spline2(grid, C, points[0], points[1], _mode_of(extrap_mode))
    else:
        if nd == 3:
            def impl(grid, C, points, out=None, k=3, diff="None", extrap_mode="linear"):
                mode = _mode_of(extrap_mode)
                res = np.empty(points.shape[0])
                for q in range(points.shape[0]):
                    res[q] = _spline3(grid, C, points[q, 0], points[q, 1], points[q, 2], mode)
                return res
        else:
            def impl(grid, C, points, out=None, k=3, diff="None", extrap_mode="linear"):
                mode = _mode_of(extrap_mode)
                res = np.empty(points.shape[0])
                for q in range(points.shape[0]):
                    res[q] = _spline2(grid, C, points[q, 0], points[q, 1], mode)
                return res
    return impl


@overload(eval_linear)
def _ov_eval_linear(grid, V, points, extrap_mode=1):
    nd = len(grid)
    if points.ndim == 1:
        if nd == 3:
            def impl(grid, V, points, extrap_mode=1):
                return _linear3(grid, V, points[0], points[1], points[2], extrap_mode)
        else:
            def impl(grid, V, points, extrap_mode=1):
                return _linear2(grid, V, points[0], points[1], extrap_mode)
    else:
        if nd == 3:
            def impl(grid, V, points, extrap_mode=1):
                res = np.empty(points.shape[0])
                for q in range(points.shape[0]):
                    res[q] = _linear3(grid, V, points[q, 0], points[q, 1], points[q, 2], extrap_mode)
                return res
        else:
            def impl(grid, V, points, extrap_mode=1):
                res = np.empty(points.shape[0])
                for q in range(points.shape[0]):
                    res[q] = _linear2(grid, V, points[q, 0], points[q, 1], extrap_mode)
                return res
    return impl


@njit
def _py_eval_spline(grid, C, points, extrap_mode):
    return eval_spline(grid, C, points, None, 3, "None", extrap_mode)


@njit
def _py_eval_linear(grid, V, points, extrap_mode):
    return eval_linear(grid, V, points, extrap_mode)
