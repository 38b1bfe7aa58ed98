"""GPU parity tests: the CUDA path (through the ctypes C-ABI of libb200cs.so) against the CPU oracle
and the frozen reference goldens.

Tolerances (north_star): flow-map positions max|dx| <= 1e-8 x domain size under identical rtol/atol;
FTLE relative L2 <= 1e-6.  The position gate is evaluated over particles whose (accepted,
rejected) step counts equal the oracle's -- DOP853's accept/reject decisions are discontinuous, so
a particle whose decision flips differs at the solver's truncation level (~1e-5) no matter how
close the arithmetic is -- and the number of such particles is gated separately.

Wall particles of the double gyre are the one systematic source of differing step sequences: the
reference evaluates sin(fl(pi*2)) = -2.4e-16 there (a rounding artefact of pi*x), so its "zero"
wall-normal velocity is 1e-17 noise that feeds hinit's h = 0.01*|y|/|f|, while the GPU's sin(pi u)
with an exact argument reduction gives an exact 0.  Those particles do not move either way
(|dx| ~ 1e-12), so the gates are stated on POSITIONS over all particles, with the step-sequence
count kept as a bounded diagnostic.

For strongly stretching cases (Bickley jet T = 6, spline fields) the step sequence of some
particles is decided by rounding noise (an error estimate err << 1 feeds the step-size formula), so
no implementation can agree with another beyond that level.  The gate there is calibrated on the
STRICT build (conftest.py `strict`, csrc/dop853.cuh B200CS_STRICT): the same kernels compiled with
the reference's evaluation order, separately rounded operations and CUDA libm differ from the
oracle only by the two libms -- that distance is the measured floor, and the product build may be
at most 3x the floor on the robust statistics (mismatch count, count above 1e-8 x L, 99th
percentile) and 10x on the single largest deviation (a heavy-tail statistic), while the bulk
(median, 99th percentile) must still meet 1e-8 x L.  profiles/r2_parity_floor.json has the numbers.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def apply_mask(arr, mask):
    out = arr.copy()
    out[mask] = 0.0
    return out


@pytest.fixture(scope="module")
def nb(lib):
    import numbacs_b200 as nb
    from numbacs_b200 import _lib
    assert _lib.device_count() >= 1
    return nb


def ftle_rel_l2(ft, fto, same=None):
    """Relative L2 error of an FTLE field; with `same` (per-particle step-count equality) the
    pixels whose 5-point stencil touches a step-count-mismatched particle are left out."""
    keep = np.ones(ft.shape, bool)
    if same is not None:
        bad = ~same
        touch = bad.copy()
        touch[1:] |= bad[:-1]
        touch[:-1] |= bad[1:]
        touch[:, 1:] |= bad[:, :-1]
        touch[:, :-1] |= bad[:, 1:]
        keep = ~touch
    return float(np.linalg.norm((ft - fto)[keep]) / np.linalg.norm(fto[keep]))


def compare_flowmaps(gpu, info, ora, steps_o, L):
    """-> dict with mismatch count and error statistics relative to the domain size L."""
    d = (np.abs(gpu - ora) / np.asarray(L)).max(axis=-1)
    same = (np.asarray(info["steps"]) == steps_o).all(axis=-1)
    return {"n": d.size, "mismatch": int((~same).sum()), "n_bad": int((d > 1e-8).sum()),
            "n_bad_match": int((d[same] > 1e-8).sum()),
            "max_match": float(d[same].max()) if same.any() else 0.0,
            "max_all": float(d.max()), "p99": float(np.percentile(d, 99)),
            "median": float(np.median(d))}


# ------------------------------------------------------------------ reference goldens (float32)

def test_flowmap_grid_2D_golden(nb, golden, coords_dg, mask_dg):
    """tests/test_integration.py:55-64 of the reference, on the GPU."""
    x, y = coords_dg
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 8.0, x, y, p)
    assert fm.shape == (21, 11, 2) and fm.dtype == np.float64
    assert np.array_equal(fm.astype(np.float32), golden["ref_fm"])
    fm_m = nb.integration.flowmap_grid_2D(f, 0.0, 8.0, x, y, p, mask=mask_dg)
    assert np.array_equal(fm_m.astype(np.float32), apply_mask(golden["ref_fm"], mask_dg))
    assert np.array_equal(fm_m[~mask_dg], fm[~mask_dg])


def test_flowmap_and_flowmap_n_golden(nb, golden, coords_dg):
    """tests/test_integration.py:38-52."""
    x, y = coords_dg
    X, Y = np.meshgrid(x, y, indexing="ij")
    pts = np.column_stack((X.ravel(), Y.ravel()))
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    fm = nb.integration.flowmap(f, 0.0, 8.0, pts, p).reshape(21, 11, 2)
    assert np.array_equal(fm.astype(np.float32), golden["ref_fm"])
    fmn, t_eval = nb.integration.flowmap_n(f, 0.0, 8.0, pts, p, n=4)
    assert np.allclose(t_eval, p[0] * np.linspace(0.0, 8.0, 4))
    assert np.array_equal(fmn.reshape(21, 11, 4, 2).astype(np.float32), golden["ref_fm_n"])


def test_flowmap_n_grid_2D_golden(nb, golden, coords_dg, mask_dg):
    """tests/test_integration.py:78-89."""
    x, y = coords_dg
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    fmn, t_eval = nb.integration.flowmap_n_grid_2D(f, 0.0, 8.0, x, y, p, n=4)
    assert fmn.shape == (21, 11, 4, 2)
    assert np.allclose(t_eval, p[0] * np.linspace(0.0, 8.0, 4))
    assert np.array_equal(fmn.astype(np.float32), golden["ref_fm_n"])
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 8.0, x, y, p)
    assert np.array_equal(fmn[:, :, -1, :], fm)          # one continuous integration
    assert np.array_equal(fmn[:, :, 0, 0], np.broadcast_to(x[:, None], (21, 11)))
    fmn_m, _ = nb.integration.flowmap_n_grid_2D(f, 0.0, 8.0, x, y, p, n=4, mask=mask_dg)
    assert np.array_equal(fmn_m.astype(np.float32), apply_mask(golden["ref_fm_n"], mask_dg))


def test_ftle_golden(nb, golden, coords_dg, mask_dg):
    """tests/test_diagnostics.py:82-96."""
    x, y = coords_dg
    ftle = nb.diagnostics.ftle_grid_2D(golden["ref_fm"], 8.0, x[1], y[1])
    assert np.allclose(ftle.astype(np.float32), golden["ref_ftle"])
    ftle_m = nb.diagnostics.ftle_grid_2D(golden["ref_fm"], 8.0, x[1], y[1], mask=mask_dg)
    assert np.allclose(ftle_m.astype(np.float32), apply_mask(golden["ref_ftle"], mask_dg))


def test_ftle_real_reference_outputs(nb, golden):
    """float64 outputs of the real numbacs.diagnostics.ftle_grid_2D; tolerance: relative L2 <= 1e-6
    (north_star); observed ~1e-16 (reciprocal multiplication instead of division)."""
    T, dx, dy = golden["ftle_args"]
    for mask, key in ((None, "ftle_out"), (golden["ftle_mask"], "ftle_out_masked")):
        got = nb.diagnostics.ftle_grid_2D(golden["ftle_in"], T, dx, dy, mask=mask)
        ref = golden[key]
        assert np.linalg.norm(got - ref) <= 1e-6 * np.linalg.norm(ref)
        assert np.abs(got - ref).max() <= 1e-12
        assert np.array_equal(got == 0.0, ref == 0.0)     # border ring / masked / lambda<=1 zeros
    out = nb.diagnostics.ftle_grid_2D(golden["ftle_in_contract"], 3.0, 1.0 / 15, 1.0 / 11)
    assert not out.any()


def test_lavd_golden(nb, golden, coords_dg, mask_dg):
    """tests/test_diagnostics.py:153-173 (trilinear vorticity, odd number of intervals)."""
    x, y = coords_dg
    X, Y = np.meshgrid(x, y, indexing="ij")
    vort = nb.flows.get_callable_scalar_linear(((0.0, 8.0, 4), (0.0, 2.0, 21), (0.0, 1.0, 11)),
                                               golden["ref_vort"])
    tspan = np.linspace(0.0, 8.0, 4)
    fmn = golden["ref_fm_n"].astype(np.float64)
    lavd = nb.diagnostics.lavd_grid_2D(fmn, tspan, 8.0, vort, X.ravel(), Y.ravel())
    assert np.allclose(lavd.astype(np.float32), golden["ref_lavd"])
    lavd_m = nb.diagnostics.lavd_grid_2D(fmn, tspan, 8.0, vort, X.ravel(), Y.ravel(), mask=mask_dg)
    assert np.allclose(lavd_m.astype(np.float32), apply_mask(golden["ref_lavd"], mask_dg))


def test_spline_tables(nb, golden):
    """tests/test_flows.py:68-139, 208-285 with the GPU prefilter / evaluators."""
    t = np.array([0.0, 0.1, 0.2])
    x = np.array([0.0, 0.5, 1.0])
    T, X, Y = np.meshgrid(t, x, x, indexing="ij")
    u = np.sin(X) * np.cos(Y) + np.sin(T)
    v = np.cos(X) * np.sin(Y) + np.cos(T)
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(t, x, x, u, v)
    assert np.allclose(grid, ((0.0, 0.2, 3), (0.0, 1.0, 3), (0.0, 1.0, 3)))
    assert Cu.shape == (5, 5, 5)
    assert np.allclose(Cu, golden["spline_Cu"]) and np.allclose(Cv, golden["spline_Cv"])
    xi = np.array([0.1, 0.4, 0.7])
    Ti, Xi, Yi = np.meshgrid(np.array([0.05, 0.12, 0.18]), xi, xi, indexing="ij")
    pts = np.column_stack((Ti.ravel(), Xi.ravel(), Yi.ravel()))
    assert np.allclose(nb.flows.get_callable_scalar(grid, Cu)(pts), golden["spline_eval_u"])
    assert np.allclose(nb.flows.get_callable_scalar(grid, Cv)(pts), golden["spline_eval_v"])
    assert np.allclose(nb.flows.get_callable_scalar_linear(grid, u)(pts), golden["linear_eval_u"])
    f1 = nb.flows.get_callable_scalar(grid, Cu)
    assert abs(f1(pts[5]) - golden["spline_eval_u"][5]) < 1e-8     # single-point call form
    gs, Cf = nb.flows.get_interp_arrays_scalar(t[::-1], x, x, u[::-1].copy())
    assert np.allclose(Cf, golden["spline_Cu"])


# ------------------------------------------------------------------ RHS level (ulp-scale)

def _gpu_rhs(lib, h, t, y, p):
    from numbacs_b200 import _lib
    t, y, p = (np.ascontiguousarray(a, dtype=np.float64) for a in (t, y, p))
    dy = np.empty_like(y)
    _lib.check(lib.b200cs_flow_rhs(h, t.ctypes.data, y.ctypes.data, len(t), p.ctypes.data, len(p),
                                   dy.ctypes.data, None))
    return dy


@pytest.mark.parametrize("name,lo,hi", [("double_gyre", (0, 0), (2, 1)),
                                        ("bickley_jet", (0, -3), (20, 3)),
                                        ("abc", (0, 0, 0), (6.28, 6.28, 6.28))])
def test_rhs_matches_oracle(nb, lib, oracle, name, lo, hi):
    """Device RHS vs the oracle's (glibc) RHS: a few ulp of the velocity scale."""
    rng = np.random.default_rng(7)
    for direction in (1.0, -1.0):
        h, p, _ = nb.flows.get_predefined_flow(name, int_direction=direction)
        ho, po, _ = oracle.get_predefined_flow(name, int_direction=direction)
        assert np.array_equal(p, po)
        y = rng.uniform(lo, hi, size=(4000, len(lo)))
        t = rng.uniform(-10, 10, size=4000)
        g = _gpu_rhs(lib, h, t, y, p)
        o = np.array([ho.rhs(t[i], y[i], po) for i in range(len(t))])
        assert np.abs(g - o).max() <= 1e-14 * np.abs(o).max()


def test_rhs_velocity_tables(nb, lib, golden):
    """Device RHS vs the reference's literal velocity tables (tests/test_flows.py:307-425)."""
    xi = np.array([0.1, 0.4, 0.7])
    Ti, Xi, Yi = np.meshgrid(np.array([0.05, 0.12, 0.18]), xi, xi, indexing="ij")
    t, y = Ti.ravel(), np.column_stack((Xi.ravel(), Yi.ravel()))
    for name, key in (("double_gyre", "dg"), ("bickley_jet", "bickley")):
        h, p, _ = nb.flows.get_predefined_flow(name)
        vel = _gpu_rhs(lib, h, t, y, p)
        assert np.allclose(vel[:, 0], golden[f"vel_{key}_u"])
        assert np.allclose(vel[:, 1], golden[f"vel_{key}_v"])
    h, p, _ = nb.flows.get_predefined_flow("abc")
    g = np.array([0.0, 0.5, 1.0]) + 0.1
    Ti, Xi, Yi, Zi = np.meshgrid(np.array([0.0, 0.1, 0.2]) + 0.1, g, g, g, indexing="ij")
    vel = _gpu_rhs(lib, h, Ti.ravel(), np.column_stack((Xi.ravel(), Yi.ravel(), Zi.ravel())), p)
    for i, c in enumerate("uvw"):
        assert np.allclose(vel[:, i], golden[f"vel_abc_{c}"])


# ------------------------------------------------------------------ flow maps vs the oracle

def test_double_gyre_C1(nb, oracle):
    """BASELINE config 1: DG 401x201, t0=0, T=-10, dop853 (README example)."""
    x, y = np.linspace(0, 2, 401), np.linspace(0, 1, 201)
    f, p, dom = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    fo, po, _ = oracle.get_predefined_flow("double_gyre", int_direction=-1.0)
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, -10.0, x, y, p, info=info)
    fmo, _, st_o, steps_o, stats_o = oracle.flowmap_grid_2D(fo, 0.0, -10.0, x, y, po, full=True)
    r = compare_flowmaps(fm, info, fmo, steps_o, (2.0, 1.0))
    assert (info["status"] == 1).all() and (st_o == 1).all()
    assert r["max_all"] <= 1e-8, r                        # north_star: 1e-8 x domain size, EVERY particle
    assert r["mismatch"] <= 4, r                          # at most the stationary corner particles
    if r["mismatch"] == 0:
        assert np.array_equal(info["stats"], stats_o)     # same nfev / accepted / rejected totals
    # FTLE from both flow maps: relative L2 <= 1e-6
    dx, dy = x[1] - x[0], y[1] - y[0]
    ft = nb.diagnostics.ftle_grid_2D(fm, -10.0, dx, dy)
    fto = oracle.ftle_grid_2D(fmo, -10.0, dx, dy)
    assert np.linalg.norm(ft - fto) <= 1e-6 * np.linalg.norm(fto)
    # fused call == the two separate calls
    fm2, ft2 = nb.diagnostics.flowmap_ftle_grid_2D(f, 0.0, -10.0, x, y, p, dx, dy)
    assert np.array_equal(fm2, fm) and np.array_equal(ft2, ft)


def test_double_gyre_damped(nb, lib, oracle):
    """alpha != 0 (flows.py:1157-1158, the -alpha*y terms) and psi != 0 run the DoubleGyreDamped
    instantiation: RHS at a few ulp of the velocity scale, flow map within 1e-8 x domain."""
    rng = np.random.default_rng(11)
    for direction in (1.0, -1.0):
        f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=direction)
        fo, po, _ = oracle.get_predefined_flow("double_gyre", int_direction=direction)
        p = p.copy()
        p[3], p[5] = 0.07, 0.3
        yy = rng.uniform((0, 0), (2, 1), size=(2000, 2))
        tt = rng.uniform(-10, 10, size=2000)
        g = _gpu_rhs(lib, f, tt, yy, p)
        o = np.array([fo.rhs(tt[i], yy[i], p) for i in range(len(tt))])
        assert np.abs(g - o).max() <= 1e-14 * np.abs(o).max()
        x, y = np.linspace(0, 2, 101), np.linspace(0, 1, 51)
        info = {}
        fm = nb.integration.flowmap_grid_2D(f, 0.0, direction * 6.0, x, y, p, info=info)
        fmo, _, st_o, steps_o, _ = oracle.flowmap_grid_2D(fo, 0.0, direction * 6.0, x, y, p, full=True)
        r = compare_flowmaps(fm, info, fmo, steps_o, (2.0, 1.0))
        assert (info["status"] == 1).all()
        assert r["max_match"] <= 1e-8 and r["mismatch"] <= 1e-3 * r["n"] + 4, r
        assert r["p99"] <= 1e-8, r
        # the undamped instantiation is the alpha -> 0 limit of the damped one
        p0 = p.copy()
        p0[3] = 0.0
        pe = p.copy()
        pe[3] = 1e-300
        a = nb.integration.flowmap_grid_2D(f, 0.0, direction * 6.0, x, y, p0)
        b = nb.integration.flowmap_grid_2D(f, 0.0, direction * 6.0, x, y, pe)
        assert np.abs(a - b).max() <= 1e-9


def assert_within_floor(r, r_strict):
    """Product-vs-oracle statistics `r` against the strict build's `r_strict` (the floor)."""
    assert r["mismatch"] <= 3 * max(r_strict["mismatch"], 1), (r, r_strict)
    assert r["n_bad_match"] <= 3 * max(r_strict["n_bad_match"], 1), (r, r_strict)
    assert r["p99"] <= max(1e-9, 3 * r_strict["p99"]), (r, r_strict)
    assert r["max_match"] <= max(1e-8, 10 * r_strict["max_match"]), (r, r_strict)


def test_bickley_jet(nb, oracle, strict):
    """BASELINE config 2 (reduced grid): Bickley jet, forward T = 6 (plot_bickley_ftle.py:24)."""
    f, p, dom = nb.flows.get_predefined_flow("bickley_jet")
    fo, po, _ = oracle.get_predefined_flow("bickley_jet")
    L = (dom[0][1] - dom[0][0], dom[1][1] - dom[1][0])
    x, y = np.linspace(dom[0][0], dom[0][1], 401), np.linspace(-3, 3, 121)
    # short horizon: plain gate
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 1.0, x, y, p, info=info)
    fmo, _, _, steps_o, _ = oracle.flowmap_grid_2D(fo, 0.0, 1.0, x, y, po, full=True)
    r = compare_flowmaps(fm, info, fmo, steps_o, L)
    assert r["mismatch"] == 0 and r["max_match"] <= 1e-8, r
    # T = 6: calibrated gate
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 6.0, x, y, p, info=info)
    fmo, _, st_o, steps_o, _ = oracle.flowmap_grid_2D(fo, 0.0, 6.0, x, y, po, full=True)
    r = compare_flowmaps(fm, info, fmo, steps_o, L)
    with strict():
        fs, ps, _ = nb.flows.get_predefined_flow("bickley_jet")
        info_s = {}
        fm_s = nb.integration.flowmap_grid_2D(fs, 0.0, 6.0, x, y, ps, info=info_s)
    r_s = compare_flowmaps(fm_s, info_s, fmo, steps_o, L)
    assert (info["status"] == 1).all()
    assert r["median"] <= 1e-12 and r["p99"] <= 1e-8, r
    assert_within_floor(r, r_s)
    dx, dy = x[1] - x[0], y[1] - y[0]
    ft = nb.diagnostics.ftle_grid_2D(fm, 6.0, dx, dy)
    fto = oracle.ftle_grid_2D(fmo, 6.0, dx, dy)
    same = (info["steps"] == steps_o).all(axis=-1)
    assert ftle_rel_l2(ft, fto, same) <= 1e-6


def test_abc_points(nb, oracle):
    """3-D state through the point-list entry (flowmap / flowmap_n, integration.py:7-120)."""
    f, p, _ = nb.flows.get_predefined_flow("abc")
    fo, po, _ = oracle.get_predefined_flow("abc")
    pts = np.random.default_rng(1).uniform(0, 2 * np.pi, size=(3000, 3))
    info = {}
    fm = nb.integration.flowmap(f, 0.0, 2.0, pts, p, info=info)
    fmo, _, _, steps_o, _ = oracle.flowmap_pts(fo, 0.0, 2.0, pts, po, full=True)
    r = compare_flowmaps(fm, info, fmo, steps_o, 2 * np.pi)
    assert r["mismatch"] <= 1 and r["max_match"] <= 1e-8, r
    fmn, ts = nb.integration.flowmap_n(f, 0.0, 2.0, pts[:500], p, n=7)
    fmno, tso = oracle.flowmap_n(fo, 0.0, 2.0, pts[:500], po, n=7)
    assert np.array_equal(ts, tso)
    assert np.abs(fmn - fmno).max() <= 1e-8 * 2 * np.pi
    with pytest.raises(ValueError):
        nb.integration.flowmap_grid_2D(f, 0.0, 1.0, np.zeros(3), np.zeros(3), p)


def _dg_like_field(nt=21, nx=41, ny=31):
    t, x, y = np.linspace(0, 10, nt), np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    T, X, Y = np.meshgrid(t, x, y, indexing="ij")
    a = 0.25 * np.sin(0.2 * np.pi * T)
    b = 1 - 2 * a
    f = a * X ** 2 + b * X
    U = -np.pi * 0.1 * np.sin(np.pi * f) * np.cos(np.pi * Y)
    V = np.pi * 0.1 * np.cos(np.pi * f) * np.sin(np.pi * Y) * (2 * a * X + b)
    return t, x, y, U, V


@pytest.mark.parametrize("mode", ["constant", "linear", "nearest"])
def test_spline_flow(nb, oracle, strict, mode):
    """Cubic-spline velocity (get_interp_arrays_2D -> get_flow_2D -> flowmap_grid_2D), particles
    inside the data grid."""
    t, x, y, U, V = _dg_like_field()
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(t, x, y, U, V)
    grid_o, Cuo, Cvo = oracle.get_interp_arrays_2D(t, x, y, U, V)
    assert np.abs(Cu - Cuo).max() <= 1e-13 and np.abs(Cv - Cvo).max() <= 1e-13
    f = nb.flows.get_flow_2D(grid, Cu, Cv, extrap_mode=mode)
    fo = oracle.get_flow_2D(grid_o, Cuo, Cvo, extrap_mode=mode)
    xg, yg = np.linspace(0.05, 1.95, 101), np.linspace(0.05, 0.95, 51)
    params = np.array([1.0])
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 8.0, xg, yg, params, info=info)
    fmo, _, _, steps_o, _ = oracle.flowmap_grid_2D(fo, 0.0, 8.0, xg, yg, params, full=True)
    r = compare_flowmaps(fm, info, fmo, steps_o, (2.0, 1.0))
    with strict():
        f_s = nb.flows.get_flow_2D(grid, Cu, Cv, extrap_mode=mode)
        info_s = {}
        fm_s = nb.integration.flowmap_grid_2D(f_s, 0.0, 8.0, xg, yg, params, info=info_s)
    r_s = compare_flowmaps(fm_s, info_s, fmo, steps_o, (2.0, 1.0))
    assert r["median"] <= 1e-12 and r["p99"] <= 1e-9, r
    assert_within_floor(r, r_s)
    # backward in time with p[0] = -1
    info = {}
    fmb = nb.integration.flowmap_grid_2D(f, 8.0, -6.0, xg, yg, -params, info=info)
    fmbo, _, _, steps_o, _ = oracle.flowmap_grid_2D(fo, 8.0, -6.0, xg, yg, -params, full=True)
    r = compare_flowmaps(fmb, info, fmbo, steps_o, (2.0, 1.0))
    assert r["mismatch"] <= 1 and r["p99"] <= 1e-9 and r["max_match"] <= 1e-8, r


@pytest.mark.parametrize("spherical", [1, 2])
def test_spline_flow_spherical(nb, oracle, spherical):
    """MERRA-shaped (config 3, reduced): lon/lat degrees, spherical=1 ([-180,180)) and 2 ([0,360)),
    including the Python-modulo longitude wrap (flows.py:162, 205)."""
    rng = np.random.default_rng(5)
    t = np.arange(25) * 1.0
    lon = (-180.0 if spherical == 1 else 0.0) + 5.0 * np.arange(72)
    lat = -90.0 + 5.0 * np.arange(37)
    T, LO, LA = np.meshgrid(t, np.deg2rad(lon), np.deg2rad(lat), indexing="ij")
    U = np.zeros_like(T)
    V = np.zeros_like(T)
    for _ in range(4):
        k, l = rng.integers(1, 4), rng.integers(1, 3)
        ph, om = rng.uniform(0, 6.28), rng.uniform(0.05, 0.2)
        U += 40.0 * np.cos(LA) * np.sin(k * LO + om * T + ph) * np.cos(l * LA)
        V += 25.0 * np.cos(LA) * np.cos(k * LO - om * T + ph) * np.sin(2 * l * LA)
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(t, lon, lat, U, V)
    f = nb.flows.get_flow_2D(grid, Cu, Cv, spherical=spherical, extrap_mode="linear")
    fo = oracle.get_flow_2D(grid, Cu, Cv, spherical=spherical, extrap_mode="linear")
    lo0 = -100.0 if spherical == 1 else 200.0
    xg, yg = lo0 + 0.5 * np.arange(120), -5.0 + 0.5 * np.arange(80)
    params = np.array([-1.0])
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 20.0, -18.0, xg, yg, params, info=info)
    fmo, _, _, steps_o, _ = oracle.flowmap_grid_2D(fo, 20.0, -18.0, xg, yg, params, full=True)
    r = compare_flowmaps(fm, info, fmo, steps_o, (360.0, 180.0))
    assert r["mismatch"] <= 1 and r["p99"] <= 1e-9 and r["max_match"] <= 1e-8, r
    # a particle that crosses the wrap longitude is handled like the reference's % operator
    edge = 179.9 if spherical == 1 else 359.9
    pts = np.array([[edge, 10.0], [edge + 0.3, -20.0], [edge - 360.0, 30.0]])
    a = nb.integration.flowmap(f, 0.0, 12.0, pts, np.array([1.0]))
    b = oracle.flowmap(fo, 0.0, 12.0, pts, np.array([1.0]))
    assert np.abs(a - b).max() <= 1e-8 * 360.0


def test_flowmap_n_dense_output(nb, oracle):
    """n = 50 output times (default of flowmap_n_grid_2D): dense-output rows vs the oracle."""
    x, y = np.linspace(0, 2, 64), np.linspace(0, 1, 33)
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    fo, po, _ = oracle.get_predefined_flow("double_gyre")
    info = {}
    fmn, ts = nb.integration.flowmap_n_grid_2D(f, 1.0, 9.0, x, y, p, info=info)
    fmno, tso, _, steps_o, stats_o = oracle.flowmap_n_grid_2D(fo, 1.0, 9.0, x, y, po, full=True)
    assert fmn.shape == (64, 33, 50, 2) and np.array_equal(ts, tso)
    same = (info["steps"] == steps_o).all(axis=-1)
    wall = np.zeros_like(same)
    wall[[0, -1], :] = True
    wall[:, [0, -1]] = True
    assert (~same).sum() <= 4 and wall[~same].all()       # only stationary corner / wall particles
    # every particle, every output time: 1e-8 x domain size (2 x 1), the north-star gate
    assert (np.abs(fmn - fmno) / np.array([2.0, 1.0])).max() <= 1e-8
    if same.all():
        assert np.array_equal(info["stats"], stats_o)      # incl. the 3 extra RHS per dense step
    # backward with p0 = -1: tspan is the physical time (integration.py:533)
    _, pb, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    fmb, tsb = nb.integration.flowmap_n_grid_2D(f, 0.0, -5.0, x[:8], y[:8], pb, n=11)
    fmbo, tsbo = oracle.flowmap_n_grid_2D(fo, 0.0, -5.0, x[:8], y[:8], pb, n=11)
    assert np.array_equal(tsb, tsbo) and np.allclose(tsb, np.linspace(0, -5, 11))
    assert np.abs(fmb - fmbo).max() <= 1e-8


def test_lavd_cubic_spline(nb, oracle):
    """LAVD with the cubic-spline vorticity (the example's path, plot_qge_elliptic_lcs.py:56-88),
    even and odd numbers of Simpson intervals, periodic wrapping, mask."""
    t, x, y, U, V = _dg_like_field(nt=17, nx=33, ny=25)
    rng = np.random.default_rng(2)
    T, X, Y = np.meshgrid(t, x, y, indexing="ij")
    vort = np.sin(3 * X + 0.3 * T) * np.cos(2 * Y) + 0.1 * rng.normal(size=T.shape)
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(t, x, y, U, V)
    f = nb.flows.get_flow_2D(grid, Cu, Cv, extrap_mode="linear")
    gw, Cw = nb.flows.get_interp_arrays_scalar(t, x, y, vort)
    w = nb.flows.get_callable_scalar(gw, Cw, extrap_mode="linear")
    wo = oracle.get_callable_scalar(gw, Cw, extrap_mode="linear")
    xg, yg = np.linspace(0.1, 1.9, 40), np.linspace(0.1, 0.9, 24)
    Xg, Yg = np.meshgrid(xg, yg, indexing="ij")
    mask = rng.random((40, 24)) < 0.1
    for n in (12, 13):
        fmn, ts = nb.integration.flowmap_n_grid_2D(f, 1.0, 6.0, xg, yg, np.array([1.0]), n=n)
        for px, py in ((0.0, 0.0), (1.5, 0.0), (0.0, 0.7), (1.5, 0.7)):
            got = nb.diagnostics.lavd_grid_2D(fmn, ts, 6.0, w, Xg.ravel(), Yg.ravel(), px, py, mask=mask)
            ref = oracle.lavd_grid_2D(fmn, ts, 6.0, wo, Xg.ravel(), Yg.ravel(), px, py, mask=mask)
            assert np.array_equal(got == 0.0, ref == 0.0)
            assert np.abs(got - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    with pytest.raises(NotImplementedError):
        nb.diagnostics.lavd_grid_2D(fmn, ts, 6.0, lambda p: p, Xg.ravel(), Yg.ravel())
    # fused path: LAVD accumulated along the trajectories inside the integration kernel
    fo = oracle.get_flow_2D(grid, Cu, Cv, extrap_mode="linear")
    for n in (12, 13):
        for px, py in ((0.0, 0.0), (1.5, 0.7)):
            got, ts, fm_end = nb.diagnostics.lavd_flowmap_grid_2D(f, 1.0, 6.0, xg, yg, np.array([1.0]), w, n=n,
                                                                  period_x=px, period_y=py, mask=mask,
                                                                  return_flowmap=True)
            fmno, tso = oracle.flowmap_n_grid_2D(fo, 1.0, 6.0, xg, yg, np.array([1.0]), n=n)
            ref = oracle.lavd_grid_2D(fmno, tso, 6.0, wo, Xg.ravel(), Yg.ravel(), px, py, mask=mask)
            assert np.array_equal(ts, tso)
            assert np.array_equal(got == 0.0, ref == 0.0)
            # trajectories agree to ~1e-9 (spline flow), so does the integral
            assert np.abs(got - ref).max() <= 1e-7 * max(1.0, np.abs(ref).max())
            assert np.abs(fm_end[~mask] - fmno[:, :, -1][~mask]).max() <= 1e-7


@pytest.mark.parametrize("linear", [False, True])
@pytest.mark.parametrize("mode", ["constant", "linear", "nearest"])
def test_lavd_time_collapsed_vorticity(nb, mode, linear):
    """The LAVD kernels evaluate the vorticity on slabs contracted over time at the output times (16 taps
    instead of 64; spatial means of an 'ij' grid through per-axis weight sums).  That re-associates the
    reference's t -> x -> y sum (flows.py:387-415 / 601-640 -> eval_spline / eval_linear): it must agree
    with the plain 3-D evaluator (B200CS_LAVD_NO_SLABS=1) to rounding, for every extrapolation mode,
    with particles and output times outside the vorticity grid, for the fused and the two-call path."""
    import os
    t, x, y, U, V = _dg_like_field(nt=17, nx=33, ny=25)
    rng = np.random.default_rng(7)
    T, X, Y = np.meshgrid(t, x, y, indexing="ij")
    vort = np.sin(3 * X + 0.3 * T) * np.cos(2 * Y) + 0.1 * rng.normal(size=T.shape)
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(t, x, y, U, V)
    f = nb.flows.get_flow_2D(grid, Cu, Cv, extrap_mode="linear")
    # the vorticity grid is SMALLER than the particle domain and the time span: the out-of-grid rules matter
    tv, xv, yv = t[2:-3], x[4:-5], y[3:-4]
    vv = vort[2:-3, 4:-5, 3:-4]
    if linear:
        w = nb.flows.get_callable_scalar_linear(((tv[0], tv[-1], len(tv)), (xv[0], xv[-1], len(xv)),
                                                 (yv[0], yv[-1], len(yv))), vv, extrap_mode=mode)
    else:
        gw, Cw = nb.flows.get_interp_arrays_scalar(tv, xv, yv, vv)
        w = nb.flows.get_callable_scalar(gw, Cw, extrap_mode=mode)
    xg, yg = np.linspace(0.1, 1.9, 40), np.linspace(0.1, 0.9, 24)
    Xg, Yg = np.meshgrid(xg, yg, indexing="ij")
    mask = rng.random((40, 24)) < 0.1
    one = np.array([1.0])
    res = {}
    for slabs in (True, False):
        if slabs:
            os.environ.pop("B200CS_LAVD_NO_SLABS", None)
        else:
            os.environ["B200CS_LAVD_NO_SLABS"] = "1"
        try:
            fused, ts = nb.diagnostics.lavd_flowmap_grid_2D(f, 1.0, 6.0, xg, yg, one, w, n=23, period_x=1.5,
                                                            mask=mask)
            fmn, ts2 = nb.integration.flowmap_n_grid_2D(f, 1.0, 6.0, xg, yg, one, n=23)
            two = nb.diagnostics.lavd_grid_2D(fmn, ts2, 6.0, w, Xg.ravel(), Yg.ravel(), 1.5, 0.0, mask=mask)
            sums = np.asarray(nb.diagnostics.lavd_vort_sums(w, ts2, Xg.ravel(), Yg.ravel()))
        finally:
            os.environ.pop("B200CS_LAVD_NO_SLABS", None)
        res[slabs] = (np.asarray(fused), np.asarray(two), sums)
    scale = max(1.0, np.abs(res[False][0]).max())
    for a, b in zip(res[True], res[False]):
        assert np.array_equal(a == 0.0, b == 0.0)
        assert np.abs(a - b).max() <= 1e-12 * max(scale, np.abs(b).max())
    assert np.abs(res[False][0]).max() > 0.0


# ------------------------------------------------------------------ API behaviour / edge cases

def test_unknown_funcptr_is_rejected(nb):
    with pytest.raises(NotImplementedError, match="no CPU fallback"):
        nb.integration.flowmap_grid_2D(140234567, 0.0, 1.0, np.zeros(2), np.zeros(2), np.ones(6))


def test_empty_and_degenerate_inputs(nb):
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 1.0, np.zeros(0), np.linspace(0, 1, 5), p)
    assert fm.shape == (0, 5, 2)
    assert nb.integration.flowmap(f, 0.0, 1.0, np.zeros((0, 2)), p).shape == (0, 2)
    # T = 0: the flow map is the identity
    x, y = np.linspace(0, 2, 7), np.linspace(0, 1, 5)
    fm = nb.integration.flowmap_grid_2D(f, 3.0, 0.0, x, y, p)
    X, Y = np.meshgrid(x, y, indexing="ij")
    assert np.array_equal(fm, np.stack([X, Y], axis=-1))
    # all masked
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 1.0, x, y, p, mask=np.ones((7, 5), bool))
    assert not fm.any()
    # tiny FTLE grids are all border
    assert not nb.diagnostics.ftle_grid_2D(np.random.rand(2, 2, 2), 1.0, 0.1, 0.1).any()
    assert nb.diagnostics.ftle_grid_2D(np.random.rand(1, 6, 2), 1.0, 0.1, 0.1).shape == (1, 6)
    # ragged sizes that are not multiples of the block / tile sizes
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 2.0, np.linspace(0, 2, 131), np.linspace(0, 1, 67), p, info=info)
    assert fm.shape == (131, 67, 2) and (info["status"] == 1).all()
    ft = nb.diagnostics.ftle_grid_2D(fm, 2.0, 2 / 130, 1 / 66)
    assert ft.shape == (131, 67) and not ft[0].any() and not ft[:, -1].any() and ft[1:-1, 1:-1].any()


def test_nan_particles_terminate_and_do_not_disturb_their_neighbours(nb):
    """A NaN initial condition must end with a non-OK status after a bounded number of rejected
    attempts (the NaN error norm rejects with h/3 until the step size is below round-off, as in
    Hairer's code) and must not change any other particle of the launch."""
    import time
    pts = np.random.default_rng(5).uniform((0, 0), (2, 1), size=(4096, 2))
    bad = pts.copy()
    bad[[7, 1000, 4095], 0] = np.nan
    bad[2048, 1] = np.nan
    isbad = np.isnan(bad).any(axis=1)
    for name in ("double_gyre", "bickley_jet"):
        f, p, _ = nb.flows.get_predefined_flow(name)
        ref = nb.integration.flowmap(f, 0.5, 4.0, pts, p)
        info = {}
        t0 = time.perf_counter()
        out = nb.integration.flowmap(f, 0.5, 4.0, bad, p, info=info)
        assert time.perf_counter() - t0 < 20.0
        assert (info["status"][isbad] != 1).all() and (info["status"][~isbad] == 1).all()
        assert np.isnan(out[isbad]).any(axis=1).all()
        assert np.array_equal(out[~isbad], ref[~isbad])


def test_torch_tensors_stay_on_device(nb):
    torch = pytest.importorskip("torch")
    f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    x = torch.linspace(0, 2, 96, dtype=torch.float64, device="cuda")
    y = torch.linspace(0, 1, 48, dtype=torch.float64, device="cuda")
    fm = nb.integration.flowmap_grid_2D(f, 0.0, -4.0, x, y, p)
    assert isinstance(fm, torch.Tensor) and fm.is_cuda and fm.shape == (96, 48, 2)
    ft = nb.diagnostics.ftle_grid_2D(fm, -4.0, 2 / 95, 1 / 47)
    assert ft.is_cuda
    fm_h = nb.integration.flowmap_grid_2D(f, 0.0, -4.0, x.cpu().numpy(), y.cpu().numpy(), p)
    assert np.array_equal(fm.cpu().numpy(), fm_h)            # same kernel, same bits
    ft_h = nb.diagnostics.ftle_grid_2D(fm_h, -4.0, 2 / 95, 1 / 47)
    assert np.array_equal(ft.cpu().numpy(), ft_h)
    # a strided / offset view still works (falls back to 8-byte stores / gets copied)
    fm2 = nb.integration.flowmap_grid_2D(f, 0.0, -4.0, x[1:], y, p)
    assert np.array_equal(fm2.cpu().numpy(), fm_h[1:])


def test_large_grid_properties(nb):
    """Size-independent properties at a size the CPU oracle would need minutes for (2048 x 1024):
    invariance of the closed DG domain, per-particle determinism (a sub-sampled grid and a point
    list give the same bits as the full grid), forward/backward round trip, FTLE border zeros."""
    torch = pytest.importorskip("torch")
    nx, ny = 2048, 1024
    f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, -10.0, x, y, p, info=info)
    assert (info["status"] == 1).all()
    # the walls are invariant for the exact flow; the solver's own truncation error (~1e-4 at
    # rtol 1e-6) lets boundary particles drift by that much
    assert fm[..., 0].min() >= -1e-3 and fm[..., 0].max() <= 2 + 1e-3
    assert fm[..., 1].min() >= -1e-3 and fm[..., 1].max() <= 1 + 1e-3
    sub = nb.integration.flowmap_grid_2D(f, 0.0, -10.0, x[::64], y[::32], p)
    assert np.array_equal(sub, fm[::64, ::32])
    X, Y = np.meshgrid(x[5::128], y[3::64], indexing="ij")
    pts = np.column_stack((X.ravel(), Y.ravel()))
    assert np.array_equal(nb.integration.flowmap(f, 0.0, -10.0, pts, p).reshape(X.shape + (2,)),
                          fm[5::128, 3::64])
    # round trip with tight tolerances
    _, pf, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=1.0)
    back = nb.integration.flowmap(f, 0.0, -10.0, pts, p, rtol=1e-11, atol=1e-13)
    again = nb.integration.flowmap(f, -10.0, 10.0, back, pf, rtol=1e-11, atol=1e-13)
    assert np.abs(again - pts).max() < 1e-6
    ft = nb.diagnostics.ftle_grid_2D(fm, -10.0, x[1] - x[0], y[1] - y[0])
    assert not ft[0].any() and not ft[-1].any() and not ft[:, 0].any() and not ft[:, -1].any()
    assert ft.min() >= 0.0 and 0.3 < ft.max() < 1.5
    # mean accepted / rejected steps as the survey measured for C1 (13.0 / 3.4)
    acc, rej = info["stats"][1] / (nx * ny), info["stats"][2] / (nx * ny)
    assert 12.0 < acc < 14.5 and 2.5 < rej < 4.5


def test_pipelined_host_path_equals_device_path(nb):
    """Large grid with host outputs: b200cs_flowmap_ftle_grid_2d integrates in row chunks and
    streams finished rows out; the result must be bit-identical to the one-shot device path."""
    torch = pytest.importorskip("torch")
    nx, ny = 2600, 1700
    f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    dx, dy = x[1] - x[0], y[1] - y[0]
    mask = np.zeros((nx, ny), bool)
    mask[100:140, 200:260] = True
    info = {}
    fm_h, ft_h = nb.diagnostics.flowmap_ftle_grid_2D(f, 0.0, -4.0, x, y, p, dx, dy, mask=mask, info=info)
    xd, yd = torch.tensor(x, device="cuda"), torch.tensor(y, device="cuda")
    fm_d = nb.integration.flowmap_grid_2D(f, 0.0, -4.0, xd, yd, p, mask=torch.tensor(mask, device="cuda"))
    ft_d = nb.diagnostics.ftle_grid_2D(fm_d, -4.0, dx, dy, mask=torch.tensor(mask, device="cuda"))
    assert np.array_equal(fm_h, fm_d.cpu().numpy())
    assert np.array_equal(ft_h, ft_d.cpu().numpy())
    assert (info["status"][~mask] == 1).all() and (info["status"][mask] == 0).all()
    assert info["stats"][1] > 0
    # FTLE only (no flow-map download) and halo rows
    _, ft_only = nb.diagnostics.flowmap_ftle_grid_2D(f, 0.0, -4.0, x, y, p, dx, dy, mask=mask, return_flowmap=False)
    assert np.array_equal(ft_only, ft_h)
    _, ft_halo = nb.diagnostics.flowmap_ftle_grid_2D(f, 0.0, -4.0, x[300:2500], y, p, dx, dy, return_flowmap=False,
                                                     halo=(1, 1))
    _, ft_full = nb.diagnostics.flowmap_ftle_grid_2D(f, 0.0, -4.0, x, y, p, dx, dy, return_flowmap=False)
    assert np.array_equal(ft_halo, ft_full[301:2499])


# ------------------------------------------------------------------ trilinear gridded flow

def test_linear_flow(nb, lib, oracle, golden):
    """get_flow_linear_2D (flows.py:418-506): RHS against the reference's eval_linear table
    (tests/test_flows.py:208-285) and the oracle, flow maps against the oracle.  A trilinear
    velocity has kinks at every cell face, where the step controller rejects and the accept /
    reject sequence is sensitive to rounding, so the flow-map gate is calibrated against the
    oracle's own one-ulp sensitivity like the spline tests."""
    # (1) the reference's literal table through the RHS entry point
    t3 = np.array([0.0, 0.1, 0.2])
    x3 = np.array([0.0, 0.5, 1.0])
    T, X, Y = np.meshgrid(t3, x3, x3, indexing="ij")
    u = np.sin(X) * np.cos(Y) + np.sin(T)
    v = np.cos(X) * np.sin(Y) + np.cos(T)
    grid3 = ((0.0, 0.2, 3), (0.0, 1.0, 3), (0.0, 1.0, 3))
    h = nb.flows.get_flow_linear_2D(grid3, u, v)
    xi = np.array([0.1, 0.4, 0.7])
    Ti, Xi, Yi = np.meshgrid(np.array([0.05, 0.12, 0.18]), xi, xi, indexing="ij")
    vel = _gpu_rhs(lib, h, Ti.ravel(), np.column_stack((Xi.ravel(), Yi.ravel())), np.array([1.0]))
    assert np.allclose(vel[:, 0], golden["linear_eval_u"]) and np.allclose(vel[:, 1], golden["linear_eval_v"])
    # (2) RHS vs oracle incl. the three extrapolation modes and points outside the grid
    t, x, y, U, V = _dg_like_field()
    grid = ((t[0], t[-1], len(t)), (x[0], x[-1], len(x)), (y[0], y[-1], len(y)))
    rng = np.random.default_rng(9)
    pts = rng.uniform((-0.2, -0.1), (2.2, 1.1), size=(3000, 2))
    ts = rng.uniform(-1, 11, size=3000)
    for mode in ("constant", "linear", "nearest"):
        f = nb.flows.get_flow_linear_2D(grid, U, V, extrap_mode=mode)
        fo = oracle.get_flow_linear_2D(grid, U, V, extrap_mode=mode)
        for p0 in (1.0, -1.0):
            g = _gpu_rhs(lib, f, p0 * ts, pts, np.array([p0]))
            o = np.array([fo.rhs(p0 * ts[i], pts[i], np.array([p0])) for i in range(len(ts))])
            assert np.abs(g - o).max() <= 1e-14
    # (3) flow maps.  With a piecewise-linear velocity the step sequence of practically every
    # particle is decided by rounding noise (measured: a one-ulp change of U changes 99 % of the
    # oracle's own step sequences and moves its particles by up to 5e-5 = the solver's truncation
    # error at rtol 1e-6), so no implementation can match another beyond that level; the GPU must
    # differ from the oracle by no more than the oracle differs from itself.
    f = nb.flows.get_flow_linear_2D(grid, U, V)
    fo = oracle.get_flow_linear_2D(grid, U, V)
    xg, yg = np.linspace(0.05, 1.95, 101), np.linspace(0.05, 0.95, 51)
    params = np.array([1.0])
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 8.0, xg, yg, params, info=info)
    fmo, _, st_o, steps_o, _ = oracle.flowmap_grid_2D(fo, 0.0, 8.0, xg, yg, params, full=True)
    assert (info["status"] == 1).all() and (st_o == 1).all()
    fo2 = oracle.get_flow_linear_2D(grid, U * (1 + 2.3e-16), V)
    fmo2 = oracle.flowmap_grid_2D(fo2, 0.0, 8.0, xg, yg, params)
    L = np.array([2.0, 1.0])
    d_gpu = (np.abs(fm - fmo) / L).max(-1)
    d_self = (np.abs(fmo2 - fmo) / L).max(-1)
    for q in (50, 90, 99):
        assert np.percentile(d_gpu, q) <= 3 * np.percentile(d_self, q) + 1e-12, (q, np.percentile(d_gpu, q))
    assert d_gpu.max() <= 10 * d_self.max()
    # with tight tolerances both the self-sensitivity and the GPU-oracle distance drop by ~1e3
    tight = dict(rtol=1e-10, atol=1e-12)
    fm_t = nb.integration.flowmap_grid_2D(f, 0.0, 8.0, xg, yg, params, **tight)
    fmo_t = oracle.flowmap_grid_2D(fo, 0.0, 8.0, xg, yg, params, **tight)
    fmo2_t = oracle.flowmap_grid_2D(fo2, 0.0, 8.0, xg, yg, params, **tight)
    d_gpu_t = (np.abs(fm_t - fmo_t) / L).max(-1)
    d_self_t = (np.abs(fmo2_t - fmo_t) / L).max(-1)
    for q in (50, 90, 99):
        assert np.percentile(d_gpu_t, q) <= 3 * np.percentile(d_self_t, q) + 1e-12, (q, np.percentile(d_gpu_t, q))
    assert np.median(d_gpu_t) <= 1e-2 * np.median(d_gpu) + 1e-12
    # spherical variant shares the expressions of the spline flow (flows.py:458-491)
    lon, lat = -180.0 + 5.0 * np.arange(72), -90.0 + 5.0 * np.arange(37)
    tt = np.arange(13) * 1.0
    Tm, LO, LA = np.meshgrid(tt, np.deg2rad(lon), np.deg2rad(lat), indexing="ij")
    Us, Vs = 40.0 * np.cos(LA) * np.sin(2 * LO + 0.1 * Tm), 25.0 * np.cos(LA) * np.cos(LO - 0.1 * Tm)
    gs = ((tt[0], tt[-1], 13), (lon[0], lon[-1], 72), (lat[0], lat[-1], 37))
    fs = nb.flows.get_flow_linear_2D(gs, Us, Vs, spherical=1, extrap_mode="linear")
    fso = oracle.get_flow_linear_2D(gs, Us, Vs, spherical=1, extrap_mode="linear")
    q = rng.uniform((-400.0, -60.0), (400.0, 60.0), size=(2000, 2))   # incl. longitudes that wrap
    tq = rng.uniform(0, 12, size=2000)
    g = _gpu_rhs(lib, fs, -tq, q, np.array([-1.0]))
    o = np.array([fso.rhs(-tq[i], q[i], np.array([-1.0])) for i in range(len(tq))])
    assert np.abs(g - o).max() <= 1e-13 * np.abs(o).max()


def test_flowmap_grid_ND(nb, oracle):
    """flowmap_grid_ND / flowmap_n_grid_ND (integration.py:185-246, 536-606): flattened 3-D initial
    conditions through the abc flow == the point-list form, and within tolerance of the oracle."""
    g = np.linspace(0.3, 5.9, 7)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    pts = np.column_stack((X.ravel(), Y.ravel(), Z.ravel()))
    f, p, _ = nb.flows.get_predefined_flow("abc")
    fo, po, _ = oracle.get_predefined_flow("abc")
    fm = nb.integration.flowmap_grid_ND(f, 0.0, 2.0, pts.ravel(), 3, p)
    assert fm.shape == (343, 3)
    assert np.array_equal(fm, nb.integration.flowmap(f, 0.0, 2.0, pts, p))
    assert np.abs(fm - oracle.flowmap(fo, 0.0, 2.0, pts, po)).max() <= 1e-8 * 2 * np.pi
    fmn, ts = nb.integration.flowmap_n_grid_ND(f, 0.0, 2.0, pts.ravel(), 3, p, n=5)
    assert fmn.shape == (343, 5, 3) and np.allclose(ts, np.linspace(0, 2, 5))
    assert np.array_equal(fmn[:, -1], fm) and np.array_equal(fmn[:, 0], pts)
    # 2-D flows work the same way
    f2, p2, _ = nb.flows.get_predefined_flow("double_gyre")
    q = np.column_stack((np.linspace(0.1, 1.9, 50), np.linspace(0.1, 0.9, 50)))
    assert np.array_equal(nb.integration.flowmap_grid_ND(f2, 0.0, 5.0, q.ravel(), 2, p2),
                          nb.integration.flowmap(f2, 0.0, 5.0, q, p2))
