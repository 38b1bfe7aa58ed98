"""CPU tests: the oracle against every golden vector / known-answer table the reference's own
tests hold for the hot path (frozen in tests/golden/reference_golden.npz).  This is what pins the
oracle; the GPU parity tests then compare the CUDA path with the oracle.

Mirrors /root/reference/tests/test_integration.py:38-89, test_diagnostics.py:82-96, 153-173,
test_flows.py:68-139, 208-285, 307-425, test_utils.py:221-232."""
import numpy as np


def apply_mask(arr, mask):
    out = arr.copy()
    out[mask] = 0.0
    return out


def test_flowmap_grid_2D_golden(oracle, golden, coords_dg, mask_dg):
    x, y = coords_dg
    f, p, _ = oracle.get_predefined_flow("double_gyre")
    fm = oracle.flowmap_grid_2D(f, 0.0, 8.0, x, y, p)
    # bit-exact after the float32 cast the reference's goldens were stored with
    assert np.array_equal(fm.astype(np.float32), golden["ref_fm"])
    fm_m = oracle.flowmap_grid_2D(f, 0.0, 8.0, x, y, p, mask=mask_dg)
    assert np.array_equal(fm_m.astype(np.float32), apply_mask(golden["ref_fm"], mask_dg))


def test_flowmap_pts_golden(oracle, golden, coords_dg):
    x, y = coords_dg
    X, Y = np.meshgrid(x, y, indexing="ij")
    pts = np.column_stack((X.ravel(), Y.ravel()))
    f, p, _ = oracle.get_predefined_flow("double_gyre")
    fm = oracle.flowmap(f, 0.0, 8.0, pts, p).reshape(21, 11, 2)
    assert np.array_equal(fm.astype(np.float32), golden["ref_fm"])
    fmn, t_eval = oracle.flowmap_n(f, 0.0, 8.0, pts, p, n=4)
    assert np.allclose(t_eval, p[0] * np.linspace(0.0, 8.0, 4))
    assert np.array_equal(fmn.reshape(21, 11, 4, 2).astype(np.float32), golden["ref_fm_n"])


def test_flowmap_n_grid_2D_golden(oracle, golden, coords_dg, mask_dg):
    x, y = coords_dg
    f, p, _ = oracle.get_predefined_flow("double_gyre")
    fmn, t_eval = oracle.flowmap_n_grid_2D(f, 0.0, 8.0, x, y, p, n=4)
    assert np.allclose(t_eval, p[0] * np.linspace(0.0, 8.0, 4))
    assert np.array_equal(fmn.astype(np.float32), golden["ref_fm_n"])
    # one continuous integration: the last row IS the final-time flow map
    assert np.array_equal(fmn[:, :, -1, :], oracle.flowmap_grid_2D(f, 0.0, 8.0, x, y, p))
    fmn_m, _ = oracle.flowmap_n_grid_2D(f, 0.0, 8.0, x, y, p, n=4, mask=mask_dg)
    assert np.array_equal(fmn_m.astype(np.float32), apply_mask(golden["ref_fm_n"], mask_dg))


def test_ftle_golden(oracle, golden, coords_dg, mask_dg):
    x, y = coords_dg
    ftle = oracle.ftle_grid_2D(golden["ref_fm"], 8.0, x[1], y[1])
    assert np.allclose(ftle.astype(np.float32), golden["ref_ftle"])
    ftle_m = oracle.ftle_grid_2D(golden["ref_fm"], 8.0, x[1], y[1], mask=mask_dg)
    assert np.allclose(ftle_m.astype(np.float32), apply_mask(golden["ref_ftle"], mask_dg))


def test_ftle_matches_real_reference_bitwise(oracle, golden):
    """Outputs of the real numbacs.diagnostics.ftle_grid_2D on seeded float64 inputs."""
    T, dx, dy = golden["ftle_args"]
    assert np.array_equal(oracle.ftle_grid_2D(golden["ftle_in"], T, dx, dy), golden["ftle_out"])
    assert np.array_equal(oracle.ftle_grid_2D(golden["ftle_in"], T, dx, dy, mask=golden["ftle_mask"]),
                          golden["ftle_out_masked"])
    fm2 = golden["ftle_in_contract"]
    out = oracle.ftle_grid_2D(fm2, 3.0, 1.0 / 15, 1.0 / 11)
    assert np.array_equal(out, golden["ftle_out_contract"])
    assert not out.any()  # max_eig <= 1 everywhere: exactly zero (diagnostics.py:62)


def test_lavd_golden(oracle, golden, coords_dg, mask_dg):
    x, y = coords_dg
    X, Y = np.meshgrid(x, y, indexing="ij")
    vort = oracle.get_callable_scalar_linear(((0.0, 8.0, 4), (0.0, 2.0, 21), (0.0, 1.0, 11)),
                                             golden["ref_vort"])
    tspan = np.linspace(0.0, 8.0, 4)
    lavd = oracle.lavd_grid_2D(golden["ref_fm_n"], tspan, 8.0, vort, X.ravel(), Y.ravel())
    assert np.allclose(lavd.astype(np.float32), golden["ref_lavd"])
    lavd_m = oracle.lavd_grid_2D(golden["ref_fm_n"], tspan, 8.0, vort, X.ravel(), Y.ravel(),
                                 mask=mask_dg)
    assert np.allclose(lavd_m.astype(np.float32), apply_mask(golden["ref_lavd"], mask_dg))


def test_simpson_real_reference(oracle, golden):
    a = oracle.composite_simpsons(golden["simpson_in_even"], 0.3)
    b = oracle.composite_simpsons(golden["simpson_in_odd"], 0.3)
    assert a == golden["simpson_out"][0] and b == golden["simpson_out"][1]


def _flows_fixture():
    t = np.array([0.0, 0.1, 0.2])
    x = np.array([0.0, 0.5, 1.0])
    y = np.array([0.0, 0.5, 1.0])
    T, X, Y = np.meshgrid(t, x, y, indexing="ij")
    u = np.sin(X) * np.cos(Y) + np.sin(T)
    v = np.cos(X) * np.sin(Y) + np.cos(T)
    xi = np.array([0.1, 0.4, 0.7])
    ti = np.array([0.05, 0.12, 0.18])
    Ti, Xi, Yi = np.meshgrid(ti, xi, xi, indexing="ij")
    pts = np.column_stack((Ti.ravel(), Xi.ravel(), Yi.ravel()))
    return t, x, y, u, v, pts


def test_spline_tables(oracle, golden):
    t, x, y, u, v, pts = _flows_fixture()
    grid, Cu, Cv = oracle.get_interp_arrays_2D(t, x, y, u, v)
    assert np.allclose(grid, ((0.0, 0.2, 3), (0.0, 1.0, 3), (0.0, 1.0, 3)))
    assert np.allclose(Cu, golden["spline_Cu"]) and np.allclose(Cv, golden["spline_Cv"])
    assert np.allclose(oracle.get_callable_scalar(grid, Cu)(pts), golden["spline_eval_u"])
    assert np.allclose(oracle.get_callable_scalar(grid, Cv)(pts), golden["spline_eval_v"])
    assert np.allclose(oracle.get_callable_scalar_linear(grid, u)(pts), golden["linear_eval_u"])
    assert np.allclose(oracle.get_callable_scalar_linear(grid, v)(pts), golden["linear_eval_v"])
    gs, Cf = oracle.get_interp_arrays_scalar(t[::-1], x, y, u[::-1])  # descending t is flipped
    assert np.allclose(Cf, golden["spline_Cu"])


def test_spline_interpolates_data(oracle):
    """Prefilter + eval reproduce the data at the knots (the defining property of `prefilter`)."""
    rng = np.random.default_rng(3)
    t, x, y = np.linspace(0, 1, 6), np.linspace(-1, 2, 9), np.linspace(3, 4, 7)
    f = rng.normal(size=(6, 9, 7))
    grid, C = oracle.get_interp_arrays_scalar(t, x, y, f)
    T, X, Y = np.meshgrid(t, x, y, indexing="ij")
    pts = np.column_stack((T.ravel(), X.ravel(), Y.ravel()))
    for mode in ("constant", "linear", "nearest"):
        got = oracle.get_callable_scalar(grid, C, extrap_mode=mode)(pts).reshape(f.shape)
        assert np.abs(got - f).max() < 1e-12


def test_velocity_tables(oracle, golden):
    """RHS of the predefined flows vs the reference's velocity tables (p[0] = +1)."""
    _, _, _, _, _, pts = _flows_fixture()
    for name, key in (("double_gyre", "dg"), ("bickley_jet", "bickley")):
        f, p, _ = oracle.get_predefined_flow(name)
        vel = np.array([f.rhs(q[0], q[1:], p) for q in pts])
        assert np.allclose(vel[:, 0], golden[f"vel_{key}_u"])
        assert np.allclose(vel[:, 1], golden[f"vel_{key}_v"])
    f, p, _ = oracle.get_predefined_flow("abc")
    g = np.array([0.0, 0.5, 1.0]) + 0.1
    Ti, Xi, Yi, Zi = np.meshgrid(np.array([0.0, 0.1, 0.2]) + 0.1, g, g, g, indexing="ij")
    pts4 = np.column_stack((Ti.ravel(), Xi.ravel(), Yi.ravel(), Zi.ravel()))
    vel = np.array([f.rhs(q[0], q[1:], p) for q in pts4])
    for i, c in enumerate("uvw"):
        assert np.allclose(vel[:, i], golden[f"vel_abc_{c}"])


def test_backward_time_convention(oracle):
    """T < 0 with p[0] = -1 integrates forward in solver time (userguide.rst:217-227); a forward
    then backward sweep returns to the start within the solver tolerance."""
    x, y = np.linspace(0.1, 1.9, 7), np.linspace(0.1, 0.9, 5)
    X, Y = np.meshgrid(x, y, indexing="ij")
    pts = np.column_stack((X.ravel(), Y.ravel()))
    f, pf, _ = oracle.get_predefined_flow("double_gyre", int_direction=1.0)
    _, pb, _ = oracle.get_predefined_flow("double_gyre", int_direction=-1.0)
    fwd = oracle.flowmap(f, 0.0, 4.0, pts, pf, rtol=1e-10, atol=1e-12)
    back = oracle.flowmap(f, 4.0, -4.0, fwd, pb, rtol=1e-10, atol=1e-12)
    assert np.abs(back - pts).max() < 1e-7
    # p[0] = +1 with T < 0 (negative h) must give the same trajectory
    back2 = oracle.flowmap(f, 4.0, -4.0, fwd, pf, rtol=1e-10, atol=1e-12)
    assert np.abs(back2 - pts).max() < 1e-7
