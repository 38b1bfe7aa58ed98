"""Parity at BASELINE.json sizes (configs 1-5), with the gates calibrated on the STRICT build.

Three integrations of the same particles are compared in every test:
  oracle   the CPU restatement of the reference (oracle/, glibc libm, unfused arithmetic),
  product  libb200cs.so (the fast kernels),
  strict   libb200cs_strict.so = the same sources with B200CS_STRICT (csrc/dop853.cuh): the
           reference's evaluation order, every operation rounded separately, CUDA libm.
strict-vs-oracle differs ONLY by the two libms (<= 1-2 ulp each) -- no GPU implementation can be
closer to the reference than that, so it is the measured floor.  Measured on B200
(profiles/r2_parity_floor.json, tests/perf/parity_floor.py):
  * double gyre (C1 full grid, C5 rows) and the MERRA-shaped spline flow (C3 full grid): product
    AND strict meet the north-star gate max|dx| <= 1e-8 x L on step-matching particles
    (C1 1.9e-9 vs 8.5e-11, C5 rows 8.0e-10 vs 1.2e-10, C3 1.8e-13 vs 5.4e-14), so the gate is the
    plain north-star one.
  * Bickley jet, T = 6 (C2): the STRICT build itself has 31 step-count mismatches in 57 696
    particles and 66 step-matching particles above 1e-8 x L (max 2.5e-6): the far field of the jet
    is almost uniform motion, its error estimate is rounding noise, and the step sequence is
    decided by the last bit of cosh / tanh.  No implementation can meet 1e-8 x L there; the product
    is gated at <= 3x the strict build's figures on the robust statistics (mismatch count, count
    above 1e-8 x L, 99th percentile) and <= 10x on the single largest deviation, which is a
    heavy-tail statistic (the RHS-only / integrator-only decomposition in
    profiles/r2_parity_decomp.json shows all of it comes from the RHS, none from the controller).
"""
import numpy as np
import pytest

from parity_common import (compare_flowmaps, ftle_rel_l2, bickley_grid, merra_axes, merra_field,
                           merra_particles, qge_field, c5_sample_rows)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb(lib):
    import numbacs_b200 as nb
    from numbacs_b200 import _lib
    assert _lib.device_count() >= 1
    return nb


def _grid(nb, flow_name, direction, t0, T, x, y, rows=None):
    f, p, _ = nb.flows.get_predefined_flow(flow_name, int_direction=direction)
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, t0, T, x, y, p, info=info)
    assert (np.asarray(info["status"]) == 1).all()
    st = np.asarray(info["steps"])
    return (fm, st) if rows is None else (fm[rows], st[rows])


def _three(nb, strict, run, ora, L):
    """run() under product and strict; -> (r_product, same_product, r_strict, same_strict, fm_p, fm_s)"""
    fm_p, st_p = run()
    with strict():
        fm_s, st_s = run()
    r_p, same_p = compare_flowmaps(fm_p, st_p, ora[0], ora[1], L)
    r_s, same_s = compare_flowmaps(fm_s, st_s, ora[0], ora[1], L)
    print("product", r_p)
    print("strict ", r_s)
    return r_p, same_p, r_s, same_s, fm_p, fm_s


def test_c1_double_gyre_full(nb, oracle, strict):
    """Config 1 at full size: every particle of the 401 x 201 grid, T = -10."""
    x, y = np.linspace(0, 2, 401), np.linspace(0, 1, 201)
    fo, po, _ = oracle.get_predefined_flow("double_gyre", int_direction=-1.0)
    fmo, _, _, so, _ = oracle.flowmap_grid_2D(fo, 0.0, -10.0, x, y, po, full=True)
    r_p, same_p, r_s, _, fm_p, _ = _three(nb, strict, lambda: _grid(nb, "double_gyre", -1.0, 0.0, -10.0, x, y),
                                         (fmo, so), (2.0, 1.0))
    assert r_s["max_rel_dx_matching"] <= 1e-8 and r_s["step_mismatches"] <= 1, r_s   # the floor itself
    assert r_p["max_rel_dx_matching"] <= 1e-8, r_p                                   # north-star gate
    assert r_p["max_rel_dx_all"] <= 1e-8, r_p          # the one mismatch is the stationary corner (0, 1)
    assert r_p["step_mismatches"] <= max(2, 3 * r_s["step_mismatches"]), (r_p, r_s)
    dx, dy = x[1] - x[0], y[1] - y[0]
    assert ftle_rel_l2(nb.diagnostics.ftle_grid_2D(fm_p, -10.0, dx, dy),
                       oracle.ftle_grid_2D(fmo, -10.0, dx, dy), same_p) <= 1e-6


def test_c2_bickley_at_size(nb, oracle, strict):
    """Config 2 at full size on the GPU (2001 x 601, T = +6); the oracle integrates 97 of the 2001
    rows (every 25th plus 16 contiguous ones for the FTLE stencil): 57 696 particles."""
    x, y = bickley_grid()
    L = (x[-1] - x[0], 6.0)
    rows = np.array(sorted(set(range(0, 2001, 25)) | set(range(1000, 1016))))
    fo, po, _ = oracle.get_predefined_flow("bickley_jet")
    fmo, _, _, so, _ = oracle.flowmap_grid_2D(fo, 0.0, 6.0, x[rows], y, po, full=True)
    r_p, same_p, r_s, same_s, fm_p, fm_s = _three(
        nb, strict, lambda: _grid(nb, "bickley_jet", 1.0, 0.0, 6.0, x, y, rows=rows), (fmo, so), L)
    # bulk: north-star figures
    assert r_p["median_rel_dx"] <= 1e-12 and r_p["p99_rel_dx"] <= 1e-8, r_p
    # tail: <= 3x the floor on the robust statistics, <= 10x on the single largest deviation
    assert r_p["step_mismatches"] <= 3 * max(r_s["step_mismatches"], 2), (r_p, r_s)
    assert r_p["over_1e-8_matching"] <= 3 * max(r_s["over_1e-8_matching"], 2), (r_p, r_s)
    assert r_p["p99_rel_dx"] <= max(1e-9, 3 * r_s["p99_rel_dx"]), (r_p, r_s)
    assert r_p["max_rel_dx_matching"] <= max(1e-8, 10 * r_s["max_rel_dx_matching"]), (r_p, r_s)
    # FTLE on the 16 contiguous rows (stencils touching a mismatch or a row outside the block excluded)
    k0 = int(np.searchsorted(rows, 1000))
    blk = slice(k0, k0 + 16)
    dx, dy = x[1] - x[0], y[1] - y[0]
    fto = oracle.ftle_grid_2D(fmo[blk], 6.0, dx, dy)
    e_p = ftle_rel_l2(nb.diagnostics.ftle_grid_2D(fm_p[blk], 6.0, dx, dy), fto, same_p[blk])
    e_s = ftle_rel_l2(nb.diagnostics.ftle_grid_2D(fm_s[blk], 6.0, dx, dy), fto, same_s[blk])
    print("ftle rel L2 product / strict:", e_p, e_s)
    assert e_p <= max(1e-6, 3 * e_s), (e_p, e_s)
    # short horizon on the same full-size grid: the plain north-star gate, no mismatches
    fmo1, _, _, so1, _ = oracle.flowmap_grid_2D(fo, 0.0, 1.0, x[rows], y, po, full=True)
    fm1, st1 = _grid(nb, "bickley_jet", 1.0, 0.0, 1.0, x, y, rows=rows)
    r1, _ = compare_flowmaps(fm1, st1, fmo1, so1, L)
    assert r1["step_mismatches"] == 0 and r1["max_rel_dx_all"] <= 1e-8, r1


@pytest.fixture(scope="module")
def merra(nb):
    import torch
    t, lon, lat = merra_axes()
    td, lond, latd = (torch.tensor(v, device="cuda") for v in (t, lon, lat))
    U, V = merra_field(torch, td, lond, latd)
    grid, Cu, Cv = nb.flows.get_interp_arrays_2D(t, lon, lat, U, V)
    del U, V
    torch.cuda.empty_cache()
    return grid, Cu, Cv


def test_c3_merra_spline_at_size(nb, oracle, strict, merra):
    """Config 3 at full size: 720 x 576 x 361 synthetic field (2 x 1.2 GB of coefficients),
    spherical = 1, every particle of the 676 x 251 grid, T = -72 h."""
    grid, Cu, Cv = merra
    lonf, latf = merra_particles()
    assert (len(lonf), len(latf)) == (676, 251)
    pm = np.array([-1.0])
    infos = []

    def run():
        fs = nb.flows.get_flow_2D(grid, Cu, Cv, spherical=1, extrap_mode="linear")
        info = {}
        fm = nb.integration.flowmap_grid_2D(fs, 360.0, -72.0, lonf, latf, pm, info=info)
        assert (np.asarray(info["status"]) == 1).all()
        infos.append(info)
        return fm, np.asarray(info["steps"])

    fso = oracle.get_flow_2D(grid, Cu.cpu().numpy(), Cv.cpu().numpy(), spherical=1, extrap_mode="linear")
    fmo, _, _, so, _ = oracle.flowmap_grid_2D(fso, 360.0, -72.0, lonf, latf, pm, full=True)
    r_p, same_p, r_s, _, fm_p, _ = _three(nb, strict, run, (fmo, so), (360.0, 180.0))
    assert r_p["step_mismatches"] <= max(1, 3 * r_s["step_mismatches"]), (r_p, r_s)
    assert r_p["max_rel_dx_matching"] <= 1e-8 and r_p["max_rel_dx_all"] <= 1e-8, r_p
    assert ftle_rel_l2(nb.diagnostics.ftle_grid_2D(fm_p, -72.0, 0.2, 0.2),
                       oracle.ftle_grid_2D(fmo, -72.0, 0.2, 0.2), same_p) <= 1e-6
    # no evaluation left the data grid: the unpinned extrapolation modes were never exercised
    if "out_of_grid" in infos[0]:
        assert int(infos[0]["out_of_grid"]) == 0


def test_c4_lavd_at_size(nb, oracle):
    """Config 4 at full size: fused LAVD over 1024 x 1024 particles with n = 601 output times on
    the QGE-shaped field; parity on a 32 x 32 strided sub-grid against the oracle's two-step path
    (flowmap_n_grid_2D + lavd_grid_2D) GIVEN the full grid's spatial-mean vorticity, and of that
    mean itself at a few output times."""
    import torch
    tq, xq, yq, U, V, vort = qge_field()
    gq, Cuq, Cvq = nb.flows.get_interp_arrays_2D(tq, xq, yq, U, V)
    fq = nb.flows.get_flow_2D(gq, Cuq, Cvq, extrap_mode="linear")
    gw, Cw = nb.flows.get_interp_arrays_scalar(tq, xq, yq, vort)
    w = nb.flows.get_callable_scalar(gw, Cw, extrap_mode="linear")
    xp, yp = np.linspace(0.02, 0.98, 1024), np.linspace(0.02, 1.98, 1024)
    one = np.array([1.0])
    n = 601
    lavd, tspan = nb.diagnostics.lavd_flowmap_grid_2D(fq, 0.5, 0.3, torch.tensor(xp, device="cuda"),
                                                      torch.tensor(yp, device="cuda"), one, w, n=n)
    lavd = _host(lavd)
    assert lavd.shape == (1024, 1024) and np.isfinite(lavd).all() and (lavd >= 0).all()
    # spatial means of the full grid (the only global quantity): GPU sums vs the oracle at 5 times
    X, Y = np.meshgrid(xp, yp, indexing="ij")
    sums = _host(nb.diagnostics.lavd_vort_sums(w, tspan, X.ravel(), Y.ravel()))
    vavg = sums / X.size
    wo = oracle.get_callable_scalar(gw, _host(Cw), extrap_mode="linear")
    for k in (0, 150, 300, 450, 600):
        pts = np.column_stack((np.full(X.size, tspan[k]), X.ravel(), Y.ravel()))
        ref = float(np.mean(wo(pts)))
        assert abs(vavg[k] - ref) <= 1e-12 * max(1.0, abs(ref)) + 1e-13 * np.abs(wo(pts)).max(), (k, vavg[k], ref)
    # trajectories + Simpson on a strided sub-grid, with the full grid's means
    sub = slice(None, None, 33)
    xs, ys = xp[sub], yp[sub]
    fqo = oracle.get_flow_2D(gq, _host(Cuq), _host(Cvq), extrap_mode="linear")
    fmno, tso = oracle.flowmap_n_grid_2D(fqo, 0.5, 0.3, xs, ys, one, n=n)
    assert np.array_equal(tso, tspan)
    # the oracle's lavd_grid_2D computes the mean over the points it is given; rebuild the integrand
    # with the full-grid mean instead (diagnostics.py:333-379 restated with numpy + its Simpson rule)
    vals = np.empty((len(xs), len(ys), n))
    for k in range(n):
        pts = np.column_stack((np.full(xs.size * ys.size, tspan[k]), fmno[:, :, k, 0].ravel(), fmno[:, :, k, 1].ravel()))
        vals[:, :, k] = np.abs(wo(pts) - vavg[k]).reshape(len(xs), len(ys))
    h = abs(tspan[1] - tspan[0])
    ref = np.array([[oracle.composite_simpsons(vals[i, j], h) for j in range(len(ys))] for i in range(len(xs))])
    got = lavd[sub, sub]
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print("LAVD rel L2 on the 32 x 32 sub-grid:", rel, "max abs", np.abs(got - ref).max(), "scale", np.abs(ref).max())
    assert rel <= 1e-6


def _host(a):
    return np.asarray(a.cpu() if hasattr(a, "cpu") else a)


def test_c5_double_gyre_rows_and_full_grid(nb, oracle, strict):
    """Config 5: 48 rows of the 16384 x 16384 grid (both borders, the middle, random blocks;
    786 432 particles) against the oracle, and the FULL 16384^2 launch against those rows
    (bit-identical: particles are independent, so the sampled rows pin the full launch)."""
    import torch
    n = 16384
    x, y = np.linspace(0, 2, n), np.linspace(0, 1, n)
    starts, rpb = c5_sample_rows(n, blocks=8, rows_per_block=6)
    rows = np.concatenate([np.arange(a, a + rpb) for a in starts])
    fo, po, _ = oracle.get_predefined_flow("double_gyre", int_direction=-1.0)
    fmo, _, _, so, _ = oracle.flowmap_grid_2D(fo, 0.0, -10.0, x[rows], y, po, full=True)
    r_p, same_p, r_s, _, fm_p, _ = _three(nb, strict, lambda: _grid(nb, "double_gyre", -1.0, 0.0, -10.0, x[rows], y),
                                         (fmo, so), (2.0, 1.0))
    assert r_s["max_rel_dx_matching"] <= 1e-8, r_s
    assert r_p["max_rel_dx_matching"] <= 1e-8, r_p
    assert r_p["mismatch_fraction"] <= 1e-5 and r_p["step_mismatches"] <= max(4, 3 * r_s["step_mismatches"]), (r_p, r_s)
    dx, dy = x[1] - x[0], y[1] - y[0]
    for b, a in enumerate(starts):
        blk = slice(b * rpb, (b + 1) * rpb)
        e = ftle_rel_l2(nb.diagnostics.ftle_grid_2D(fm_p[blk], -10.0, dx, dy),
                        oracle.ftle_grid_2D(fmo[blk], -10.0, dx, dy), same_p[blk])
        assert e <= 1e-6, (a, e)
    # the full-size launch
    f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    xd, yd = torch.tensor(x, device="cuda"), torch.tensor(y, device="cuda")
    full = nb.integration.flowmap_grid_2D(f, 0.0, -10.0, xd, yd, p, device_out=True)
    got = full[torch.tensor(rows, device="cuda")].cpu().numpy()
    assert np.array_equal(got, fm_p)
    ft = nb.diagnostics.ftle_grid_2D(full, -10.0, dx, dy, device_out=True)
    assert bool(torch.isfinite(ft).all()) and float(ft[0].abs().max()) == 0.0 and float(ft[:, -1].abs().max()) == 0.0
    b = starts.index(n // 2 - rpb // 2)
    mid = starts[b]
    assert np.array_equal(ft[mid + 1:mid + rpb - 1].cpu().numpy(),
                          nb.diagnostics.ftle_grid_2D(fm_p[b * rpb:(b + 1) * rpb], -10.0, dx, dy)[1:-1])
