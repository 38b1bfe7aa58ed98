"""Config 4's call chain through the drop-in (examples/elliptic_lcs/plot_qge_elliptic_lcs.py:47-88):
get_interp_arrays_2D -> get_flow_2D / get_callable_2D -> flowmap_n_grid_2D -> curl_func_tspan ->
get_interp_arrays_scalar -> get_callable_scalar -> lavd_grid_2D, against frozen outputs of the REAL
reference running over oracle/shims (tests/golden/make_config4_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config4_golden.npz"))


@pytest.fixture(scope="module")
def nb(lib):
    import numbacs_b200 as nb
    from numbacs_b200 import _lib
    assert _lib.device_count() >= 1
    return nb


def test_get_callable_2D(nb, G):
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(G["t"], G["x"], G["y"], G["U"], G["V"])
    vel = nb.flows.get_callable_2D(grid, Cu, Cv, extrap_mode="linear")
    got = vel(G["vel_pts"])
    assert got.shape == (40, 2)
    assert np.abs(got - G["vel_vals"]).max() <= 1e-14
    one = vel(G["vel_pts"][0])
    assert isinstance(one, np.ndarray) and one.shape == (2,) and np.abs(one - G["vel_vals"][0]).max() <= 1e-14
    velt = nb.flows.get_callable_2D(grid, Cu, Cv, extrap_mode="linear", return_type="tuple")
    tu = velt(G["vel_pts"][0])
    assert isinstance(tu, tuple) and np.abs(np.array(tu) - G["vel_tuple0"]).max() <= 1e-14
    vs = nb.flows.get_callable_2D(grid, Cu, Cv, spherical=1, extrap_mode="linear", r=3.5)
    assert np.abs(vs(G["vel_pts"]) - G["vel_sph_vals"]).max() <= 1e-13 * np.abs(G["vel_sph_vals"]).max()
    # spherical = 2 is NOT scaled by the reference's callable (it only tests spherical == 1)
    v2 = nb.flows.get_callable_2D(grid, Cu, Cv, spherical=2, extrap_mode="linear", r=3.5)
    assert np.array_equal(v2(G["vel_pts"]), got)
    with pytest.raises(ValueError):
        nb.flows.get_callable_2D(grid, Cu, Cv, return_type="list")


def test_config4_chain_matches_the_reference(nb, G):
    t, x, y, xg, yg, n = G["t"], G["x"], G["y"], G["xg"], G["yg"], int(G["n"][0])
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(t, x, y, G["U"], G["V"])
    funcptr = nb.flows.get_flow_2D(grid, Cu, Cv, extrap_mode="linear")
    vel = nb.flows.get_callable_2D(grid, Cu, Cv, extrap_mode="linear")
    params = np.array([1.0])
    fmn, tspan = nb.integration.flowmap_n_grid_2D(funcptr, 0.3, 0.4, xg, yg, params, n=n)
    assert np.array_equal(tspan, G["tspan"])
    assert np.abs(fmn - G["fmn"]).max() <= 1e-8 * 2.0
    vort = nb.utils.curl_func_tspan(vel, tspan, xg, yg, h=1e-3)
    assert vort.shape == G["vort"].shape
    # a central difference over 2h = 2e-3 amplifies the 1e-16 differences of the interpolant by 1e3
    assert np.abs(vort - G["vort"]).max() <= 1e-12 * np.abs(G["vort"]).max() + 1e-12
    gw, Cw = nb.flows.get_interp_arrays_scalar(tspan, xg, yg, vort)
    w = nb.flows.get_callable_scalar(gw, Cw)
    Xg, Yg = np.meshgrid(xg, yg, indexing="ij")
    lavd = nb.diagnostics.lavd_grid_2D(fmn, tspan, 0.4, w, Xg.ravel(), Yg.ravel())
    rel = np.linalg.norm(lavd - G["lavd"]) / np.linalg.norm(G["lavd"])
    print("config-4 chain: LAVD rel L2 vs the reference", rel)
    assert rel <= 1e-6
    # the fused path gives the same field
    fused, _ = nb.diagnostics.lavd_flowmap_grid_2D(funcptr, 0.3, 0.4, xg, yg, params, w, n=n)
    assert np.linalg.norm(fused - G["lavd"]) / np.linalg.norm(G["lavd"]) <= 1e-6
    # device-resident inputs / outputs
    import torch
    vd = nb.utils.curl_func_tspan(vel, torch.tensor(tspan, device="cuda"), torch.tensor(xg, device="cuda"),
                                  torch.tensor(yg, device="cuda"), h=1e-3)
    assert vd.is_cuda and np.array_equal(vd.cpu().numpy(), vort)
    with pytest.raises(NotImplementedError):
        nb.utils.curl_func_tspan(lambda p: p, tspan, xg, yg)


def test_out_of_grid_counter(nb, G):
    """info['out_of_grid'] counts right-hand-side evaluations outside the data grid (the guard on the
    extrapolation modes, which no reference test pins): 0 for particles that stay inside, > 0 once a
    particle starts outside, 0 again for the next call (the counter is reset when it is read)."""
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(G["t"], G["x"], G["y"], G["U"], G["V"])
    f = nb.flows.get_flow_2D(grid, Cu, Cv, extrap_mode="linear")
    params = np.array([1.0])
    info = {}
    nb.integration.flowmap_grid_2D(f, 0.3, 0.4, G["xg"], G["yg"], params, info=info)
    assert info["out_of_grid"] == 0
    info = {}
    nb.integration.flowmap(f, 0.3, 0.4, np.array([[0.5, 1.0], [1.5, 1.0], [0.5, -0.2]]), params, info=info)
    assert info["out_of_grid"] > 0
    assert nb.integration.out_of_grid_count(f) == 0
    info = {}
    nb.integration.flowmap(f, 0.9, 0.4, np.array([[0.5, 1.0]]), params, info=info)   # leaves the grid in TIME
    assert info["out_of_grid"] > 0
    fa, pa, _ = nb.flows.get_predefined_flow("double_gyre")
    info = {}
    nb.integration.flowmap_grid_2D(fa, 0.0, 1.0, np.linspace(0, 2, 5), np.linspace(0, 1, 4), pa, info=info)
    assert info["out_of_grid"] == 0                                                # analytic flows have no grid


def test_dense_rows_of_a_failed_integration_are_nan(nb):
    """A particle whose integration gives up (NaN state -> step size underflow) leaves NaN in the
    output rows it never reached instead of uninitialised memory; the other particles are untouched."""
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    pts = np.array([[0.5, 0.5], [np.nan, 0.3], [1.5, 0.25]])
    info = {}
    fmn, ts = nb.integration.flowmap_n(f, 0.0, 4.0, pts, p, n=6, info=info)
    st = np.asarray(info["status"])
    assert st[0] == 1 and st[2] == 1 and st[1] != 1
    assert np.isfinite(fmn[0]).all() and np.isfinite(fmn[2]).all()
    assert np.isnan(fmn[1, 1:]).all()
    ok, _ = nb.integration.flowmap_n(f, 0.0, 4.0, pts[[0, 2]], p, n=6)
    assert np.array_equal(ok, fmn[[0, 2]])
