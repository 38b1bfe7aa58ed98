"""GPU parity tests for the rows of SURVEY.md section 8(f) built so far: aux-grid flow map,
Cauchy-Green tensor / eigen-pairs, ftle_from_eig, FTLE ridge points, order statistics -- all
through the ctypes C-ABI of libb200cs.so, against the reference's goldens, the frozen outputs of
the real reference code, and the CPU oracle.

The tensor / eigen / ridge kernels use explicitly rounded arithmetic in the reference's operation
order, so they are required to be BIT-IDENTICAL to the reference outputs (eigenvector signs
included); the aux-grid flow map carries the usual solver tolerance (1e-8 x domain over particles
with identical step sequences)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def apply_mask(arr, mask):
    out = arr.copy()
    out[mask] = 0.0
    return out


def reconstruct_matrix(evals, evecs):
    return np.einsum("...ik,...k,...jk->...ij", evecs, evals, evecs)


@pytest.fixture(scope="module")
def nb(lib):
    import numbacs_b200 as nb
    from numbacs_b200 import _lib
    assert _lib.device_count() >= 1
    return nb


# ------------------------------------------------------------------ flowmap_aux_grid_2D

def test_flowmap_aux_golden(nb, golden, coords_dg, mask_dg):
    """tests/test_integration.py:66-75 of the reference, on the GPU."""
    x, y = coords_dg
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    fa = nb.integration.flowmap_aux_grid_2D(f, 0.0, 8.0, x, y, p)
    assert fa.shape == (21, 11, 5, 2) and fa.dtype == np.float64
    assert np.array_equal(fa.astype(np.float32), golden["ref_fm_aux"])
    assert np.array_equal(fa[:, :, 4, :], nb.integration.flowmap_grid_2D(f, 0.0, 8.0, x, y, p))
    fa_m = nb.integration.flowmap_aux_grid_2D(f, 0.0, 8.0, x, y, p, mask=mask_dg)
    assert np.array_equal(fa_m.astype(np.float32), apply_mask(golden["ref_fm_aux"], mask_dg))


@pytest.mark.parametrize("eig_main,compute_edge", [(True, True), (True, False), (False, True), (False, False)])
def test_flowmap_aux_vs_oracle(nb, oracle, eig_main, compute_edge):
    x, y = np.linspace(0, 2, 61), np.linspace(0, 1, 37)
    rng = np.random.default_rng(3)
    mask = rng.random((61, 37)) < 0.1
    f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    fo, po, _ = oracle.get_predefined_flow("double_gyre", int_direction=-1.0)
    info = {}
    fa = nb.integration.flowmap_aux_grid_2D(f, 0.0, -10.0, x, y, p, h=1e-5, eig_main=eig_main,
                                            compute_edge=compute_edge, mask=mask, info=info)
    fao, status_o, steps_o, stats_o = oracle.flowmap_aux_grid_2D(
        fo, 0.0, -10.0, x, y, po, h=1e-5, eig_main=eig_main, compute_edge=compute_edge, mask=mask,
        full=True)
    assert fa.shape == fao.shape
    # exactly the same set of integrated entries, zeros elsewhere
    assert np.array_equal(info["status"] == 1, status_o == 1)
    assert not fa[status_o != 1].any()
    same = (info["steps"] == steps_o).all(axis=-1)
    assert (~same).sum() <= max(1, int(1e-4 * same.size))
    d = (np.abs(fa - fao) / np.array([2.0, 1.0])).max(axis=-1)
    assert d[same].max() <= 1e-8
    assert tuple(info["stats"]) == tuple(stats_o) or (~same).any()


# ------------------------------------------------------------------ tensor / eigen-pairs

def test_tensor_goldens(nb, golden, coords_dg, mask_dg):
    """tests/test_diagnostics.py:99-150 of the reference, on the GPU."""
    x, y = coords_dg
    D = nb.diagnostics
    fa = golden["ref_fm_aux"].astype(np.float64)
    fm = golden["ref_fm"].astype(np.float64)
    C = D.C_tensor_2D(fa, x[1], y[1])
    assert np.allclose(C.astype(np.float32), golden["ref_C"])
    Cm = D.C_tensor_2D(fa, x[1], y[1], mask=mask_dg)
    assert np.allclose(Cm.astype(np.float32), apply_mask(golden["ref_C"], mask_dg))
    for fn, inp, gv, ge in ((D.C_eig_aux_2D, fa, "ref_Cevals_aux", "ref_Cevecs_aux"),
                            (D.C_eig_2D, fm, "ref_Cevals", "ref_Cevecs")):
        Cexp = reconstruct_matrix(golden[gv], golden[ge])
        vals, vecs = fn(inp, x[1], y[1])
        assert np.allclose(reconstruct_matrix(vals.astype(np.float32), vecs.astype(np.float32)), Cexp)
        vals, vecs = fn(inp, x[1], y[1], mask=mask_dg)
        assert np.allclose(reconstruct_matrix(vals.astype(np.float32), vecs.astype(np.float32)),
                           apply_mask(Cexp, mask_dg))


def test_tensor_functions_bit_identical_to_reference(nb, golden):
    """Frozen outputs of the real numbacs.diagnostics code (numba + LAPACK eigh)."""
    D = nb.diagnostics
    dx, dy, h = golden["ceig_args"]
    fm, fa, mask = golden["ceig_in"], golden["caux_in"], golden["ceig_mask"]
    for tag, m in (("", None), ("_masked", mask)):
        vals, vecs = D.C_eig_2D(fm, dx, dy, m)
        assert np.array_equal(vals, golden["ceig_vals" + tag])
        assert np.array_equal(vecs, golden["ceig_vecs" + tag])
        assert np.array_equal(D.C_tensor_2D(fa, dx, dy, h, m), golden["ctensor" + tag])
        vals, vecs = D.C_eig_aux_2D(fa, dx, dy, h, True, m)
        assert np.array_equal(vals, golden["caux_vals_main" + tag])
        assert np.array_equal(vecs, golden["caux_vecs_main" + tag])
        vals, vecs = D.C_eig_aux_2D(np.ascontiguousarray(fa[:, :, :4]), dx, dy, h, False, m)
        assert np.array_equal(vals, golden["caux_vals" + tag])
        assert np.array_equal(vecs, golden["caux_vecs" + tag])
    ft = D.ftle_from_eig(golden["ceig_vals"][:, :, 1], -3.0)        # strided view, read in place
    assert ft.shape == golden["ftle_from_eig_out"].shape
    assert np.allclose(ft, golden["ftle_from_eig_out"], rtol=1e-15, atol=0)
    assert np.array_equal(ft == 0, golden["ftle_from_eig_out"] == 0)


def test_eigh_degenerate_matrices(nb, oracle):
    """Flow maps chosen so that the Cauchy-Green tensor is diagonal, a multiple of the identity,
    zero, or has a negligible off-diagonal: the split branch of LAPACK's dsteqr."""
    X, Y = np.meshgrid(np.linspace(0, 1, 9), np.linspace(0, 1, 7), indexing="ij")
    cases = [np.stack([2 * X, 3 * Y], -1), np.stack([3 * X, 2 * Y], -1), np.stack([X, Y], -1),
             np.zeros(X.shape + (2,)), np.stack([2 * X + 1e-17 * Y, 3 * Y], -1),
             np.stack([X + 0.5 * Y, Y], -1), np.stack([Y, X], -1)]
    for fm in cases:
        vals, vecs = nb.diagnostics.C_eig_2D(fm, 0.125, 1 / 6)
        vo, eo = oracle.C_eig_2D(fm, 0.125, 1 / 6)
        assert np.array_equal(vals, vo) and np.array_equal(vecs, eo)


def test_c_eig_matches_ftle_kernel_and_oracle_on_a_real_flow_map(nb, oracle):
    x, y = np.linspace(0, 2, 301), np.linspace(0, 1, 151)
    dx, dy = x[1] - x[0], y[1] - y[0]
    f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    fm = nb.integration.flowmap_grid_2D(f, 0.0, -10.0, x, y, p)
    vals, vecs = nb.diagnostics.C_eig_2D(fm, dx, dy)
    vo, eo = oracle.C_eig_2D(fm, dx, dy)
    assert np.array_equal(vals, vo) and np.array_equal(vecs, eo)
    # unit eigenvectors, C v = lambda v reconstructs the tensor
    assert np.allclose(np.linalg.norm(vecs[1:-1, 1:-1], axis=2), 1.0, atol=1e-14)
    ft_eig = nb.diagnostics.ftle_from_eig(vals[:, :, 1], -10.0)
    ft = nb.diagnostics.ftle_grid_2D(fm, -10.0, dx, dy)
    assert np.linalg.norm(ft_eig - ft) / np.linalg.norm(ft) <= 1e-9   # same field, two routes
    assert np.allclose(ft_eig, oracle.ftle_from_eig(vo[:, :, 1], -10.0), rtol=1e-14, atol=0)
    # torch tensors stay on the device
    import torch
    fmd = torch.from_numpy(fm).cuda()
    vd, ed = nb.diagnostics.C_eig_2D(fmd, dx, dy)
    assert vd.is_cuda and ed.is_cuda
    assert np.array_equal(vd.cpu().numpy(), vals) and np.array_equal(ed.cpu().numpy(), vecs)
    ftd = nb.diagnostics.ftle_from_eig(vd[:, :, 1], -10.0)
    assert ftd.is_cuda and np.array_equal(ftd.cpu().numpy(), ft_eig)


# ------------------------------------------------------------------ ridge points

def test_ridge_pts_golden(nb, golden, coords_dg):
    """tests/test_extraction.py:7-13 of the reference, on the GPU."""
    x, y = coords_dg
    r = nb.extraction.ftle_ridge_pts(golden["ref_ftle"], golden["ref_Cevecs"][:, :, :, 1], x, y)
    assert r.shape == golden["ref_ridge_pts"].shape
    assert np.allclose(r, golden["ref_ridge_pts"])


def test_ridge_pts_bit_identical_to_reference(nb, golden):
    """Frozen outputs of the real numbacs/extraction/ridges.py on a seeded float64 field, incl.
    sdd_thresh and the percentile threshold (np.percentile as numba computes it)."""
    from numbacs_b200.extraction import _ftle_ridge_pts_connect, percentile_value
    f, ev, x, y = golden["ridge_f"], golden["ridge_ev"], golden["ridge_x"], golden["ridge_y"]
    for tag, (thr, pct) in zip("abc", golden["ridge_args"]):
        r = nb.extraction.ftle_ridge_pts(f, ev, x, y, thr, int(pct))
        assert np.array_equal(r, golden["ridge_pts_" + tag])
        rp, rv, sdd, h = _ftle_ridge_pts_connect(f, ev, x, y, thr, int(pct))
        assert np.array_equal(rp, golden["ridge_conn_pts_" + tag])
        assert np.array_equal(rv, golden["ridge_conn_vec_" + tag])
        assert np.array_equal(sdd, golden["ridge_conn_sdd_" + tag])
        assert h == min(x[1] - x[0], y[1] - y[0])
    for pct in (1, 25, 50, 60, 99.5, 100):
        assert percentile_value(f, pct) == np.percentile(f, pct) or \
            abs(percentile_value(f, pct) - np.percentile(f, pct)) <= 1e-15 * abs(np.percentile(f, pct))


def test_order_stats_against_sort(nb, lib):
    import ctypes as C
    rng = np.random.default_rng(11)
    for data in (rng.normal(size=100003), np.round(rng.normal(size=50000), 1),     # many ties
                 np.concatenate([np.zeros(3000), -np.zeros(10), rng.random(100)]),  # mostly zeros
                 np.array([3.5]), -rng.random(777) * 1e-300, rng.normal(size=4096) * 1e300):
        s = np.sort(data)
        n = data.size
        for k in sorted({0, n // 3, n // 2, n - 2, n - 1} & set(range(n))):
            out = np.zeros(2)
            rc = lib.b200cs_order_stats(C.c_void_p(data.ctypes.data), n, k,
                                        C.c_void_p(out.ctypes.data), None)
            assert rc == 0
            assert out[0] == s[k] and out[1] == s[min(k + 1, n - 1)], (n, k, out, s[k])


def test_ridge_pipeline_on_a_real_ftle_field(nb, oracle):
    """Config 5's tail at a size the oracle finishes in seconds: flow map -> C_eig_2D ->
    ftle_from_eig -> ridge points (plot_dg_ftle_ridges.py:43-72), device-resident end to end."""
    import torch
    x, y = np.linspace(0, 2, 401), np.linspace(0, 1, 201)
    dx, dy = x[1] - x[0], y[1] - y[0]
    f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    fm = nb.integration.flowmap_grid_2D(f, 0.0, -10.0, x, y, p, device_out=True)
    vals, vecs = nb.diagnostics.C_eig_2D(fm, dx, dy)
    ftle = nb.diagnostics.ftle_from_eig(vals[:, :, 1], -10.0)
    r = nb.extraction.ftle_ridge_pts(ftle, vecs[:, :, :, 1], x, y, sdd_thresh=10.0, percentile=50)
    assert r.is_cuda and r.shape[1] == 2 and r.shape[0] > 100
    # oracle on the same (GPU-produced) fields
    ftle_h, vecs_h = ftle.cpu().numpy(), vecs.cpu().numpy()
    ro = oracle.ftle_ridge_pts(ftle_h, vecs_h[:, :, :, 1], x, y, sdd_thresh=10.0, percentile=50)
    assert np.array_equal(r.cpu().numpy(), ro)
    # every ridge point lies within half a cell of a grid point with ftle above the median
    rr = r.cpu().numpy()
    i = np.rint(rr[:, 0] / dx).astype(int)
    j = np.rint(rr[:, 1] / dy).astype(int)
    assert (np.abs(rr[:, 0] - x[i]) <= dx / 2 + 1e-12).all() and (np.abs(rr[:, 1] - y[j]) <= dy / 2 + 1e-12).all()
    assert (ftle_h[i, j] > np.percentile(ftle_h, 50)).all()


def _check_ridges(ridges, cat, lens, exact):
    assert [len(r) for r in ridges] == list(lens)
    got = np.concatenate(ridges)
    assert np.array_equal(got, cat) if exact else np.allclose(got, cat)


def test_ftle_ridges_golden_and_reference(nb, golden, coords_dg):
    """tests/test_extraction.py:15-21 (ridges.pkl) and frozen outputs of the real ftle_ridges:
    GPU union-find labelling == scipy.ndimage.label grouping, label order and point order included."""
    x, y = coords_dg
    r = nb.extraction.ftle_ridges(golden["ref_ftle"], golden["ref_Cevecs"][:, :, :, 1], x, y)
    _check_ridges(r, golden["ref_ridges_cat"], golden["ref_ridges_len"], exact=False)
    f, ev, xr, yr = golden["ridge_f"], golden["ridge_ev"], golden["ridge_x"], golden["ridge_y"]
    for tag, (thr, pct, mrp) in zip("ab", golden["ridges_args"]):
        r = nb.extraction.ftle_ridges(f, ev, xr, yr, thr, int(pct), int(mrp))
        _check_ridges(r, golden["ridges_cat_" + tag], golden["ridges_len_" + tag], exact=True)


def test_ftle_ridges_long_winding_components(nb, oracle):
    """Components that wind across many thread blocks (spirals, a comb, diagonal staircases):
    the union-find must agree with scipy's labelling whatever the merge order."""
    n = 257
    x, y = np.linspace(0, 1, n), np.linspace(0, 1, n + 30)
    X, Y = np.meshgrid(x, y, indexing="ij")
    R, TH = np.hypot(X - 0.5, Y - 0.5), np.arctan2(Y - 0.5, X - 0.5)
    rng = np.random.default_rng(5)
    fields = [1.0 + np.cos(40 * R - 3 * TH),                       # three-armed spiral ridges
              1.0 + np.cos(60 * X) * (Y > 0.1) + np.cos(60 * Y) * (Y <= 0.1),   # comb joined by a spine
              1.0 + np.cos(50 * (X + Y)) + 0.05 * rng.normal(size=X.shape)]     # noisy diagonals
    for f in fields:
        gx, gy = np.gradient(f, x, y)
        nrm = np.hypot(gx, gy) + 1e-300
        ev = np.stack([gx / nrm, gy / nrm], -1)                    # ridge normal ~ gradient direction
        r = nb.extraction.ftle_ridges(f, ev, x, y, 0.0, 0, 1)
        ro = oracle.ftle_ridges(f, ev, x, y, 0.0, 0, 1)
        assert len(r) == len(ro) and len(r) > 0
        assert [len(a) for a in r] == [len(a) for a in ro]
        assert np.array_equal(np.concatenate(r), np.concatenate(ro))


# ------------------------------------------------------------------ flow-map composition

def test_flowmap_composition_golden(nb, golden, coords_dg):
    """tests/test_integration.py:92-114 of the reference (fm_ci / fms_ci / fm_cs / fms_cs.npy)."""
    from test_oracle_tensor_golden import check_composition_initial
    I = nb.integration
    x, y = coords_dg
    grid = ((x[0], x[-1], 21), (y[0], y[-1], 11))
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    fm0, fms, nT = I.flowmap_composition_initial(f, 0.0, 8.0, 1.0, x, y, grid, p)
    assert fm0.shape == (21, 11, 2) and fms.shape == (8, 21, 11, 2)
    check_composition_initial(fm0, fms, nT, golden)
    fms_in = golden["ref_fms_ci"].astype(np.float64)
    fmk, fms2 = I.flowmap_composition_step(fms_in, f, 8.0, 1.0, 8, x, y, grid, p)
    assert np.allclose(fmk.astype(np.float32), golden["ref_fm_cs"])
    assert np.allclose(fms2.astype(np.float32), golden["ref_fms_cs"])
    assert fms2 is fms_in                                   # updated in place, like the reference


def test_flowmap_composition_vs_oracle_device_resident(nb, oracle):
    import torch
    I = nb.integration
    nx, ny, nT = 301, 151, 10
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    grid = ((x[0], x[-1], nx), (y[0], y[-1], ny))
    f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=-1.0)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    fm0, fms, n = I.flowmap_composition_initial(f, 0.0, -10.0, -1.0, xd, yd, grid, p)
    assert n == nT and fm0.is_cuda and fms.is_cuda and fms.shape == (nT, nx, ny, 2)
    # each intermediate map is the plain flow map over its own interval
    k = 3
    assert np.array_equal(fms[k].cpu().numpy(), I.flowmap_grid_2D(f, -3.0, -1.0, x, y, p))
    # oracle composition of the SAME intermediate maps; the walls are left out (positions there
    # sit within an ulp of the grid edge, where CONSTANT extrapolation switches to 0)
    ref = oracle.flowmap_composition(fms.cpu().numpy(), grid, nT)
    got = fm0.cpu().numpy()
    assert np.abs(got - ref)[1:-1, 1:-1].max() <= 1e-12     # 1e-16 x the flow-map gradient (<= 1e3)
    # step: same as recomputing from scratch one interval later
    fmk, fms = I.flowmap_composition_step(fms, f, -10.0, -1.0, nT, xd, yd, grid, p)
    fm0b, fmsb, _ = I.flowmap_composition_initial(f, -1.0, -10.0, -1.0, xd, yd, grid, p)
    assert torch.equal(fms, fmsb) and torch.equal(fmk, fm0b)
    # the composed map approximates the directly integrated one
    direct = I.flowmap_grid_2D(f, 0.0, -10.0, x, y, p)
    assert np.median(np.abs(got - direct)) < 2e-3
    # a point that leaves the grid gets 0 (CONSTANT extrapolation); nT = 1 applies the map to itself
    fms1 = fms[:1].clone()
    one = I.flowmap_composition(fms1, grid, 1)
    assert np.abs(one.cpu().numpy() - oracle.flowmap_composition(fms1.cpu().numpy(), grid, 1))[1:-1, 1:-1].max() <= 1e-12
    fms[0, 5, 7, 0] = 2.5
    out = I.flowmap_composition(fms, grid, nT)
    assert not out[5, 7].any()


# ------------------------------------------------------------------ mask helpers

def test_mask_helpers(nb, golden, oracle):
    import torch
    U = nb.utils
    m = golden["dil_in"]
    d4, d8 = U.binary_mask_dilation(m), U.binary_mask_dilation(m, corners=True)
    assert d4.dtype == np.bool_ and np.array_equal(d4, golden["dil_out4"]) and np.array_equal(d8, golden["dil_out8"])
    md = torch.from_numpy(m).cuda()
    assert torch.equal(U.binary_mask_dilation(md, corners=True).cpu(), torch.from_numpy(golden["dil_out8"]))
    rng = np.random.default_rng(2)
    big = rng.random((515, 1027)) < 0.01
    assert np.array_equal(U.binary_mask_dilation(big), oracle.binary_mask_dilation(big))
    for shape in ((1, 7), (7, 1), (1, 1)):
        one = np.zeros(shape, bool)
        one[0, 0] = True
        assert np.array_equal(U.binary_mask_dilation(one, corners=True), oracle.binary_mask_dilation(one, True))
    # fill_nans_and_get_mask: in place, mask from the first time slice
    u = rng.normal(size=(3, 5, 4))
    v = rng.normal(size=(3, 5, 4))
    u[:, 1, 2] = np.nan
    v[:, 1, 2] = np.nan
    u2, v2, mask = U.fill_nans_and_get_mask((u, v))
    assert u2 is u and mask.shape == (5, 4) and mask.sum() == 1 and mask[1, 2]
    assert not np.isnan(u).any() and (u[:, 1, 2] == 0).all() and (v[:, 1, 2] == 0).all()
    # the dilated mask keeps the C_eig stencil away from masked flow-map entries
    x, y = np.linspace(0, 2, 41), np.linspace(0, 1, 21)
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    hole = np.zeros((41, 21), bool)
    hole[10:13, 5:8] = True
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 5.0, x, y, p, mask=hole)
    vals, _ = nb.diagnostics.C_eig_2D(fm, x[1], y[1], U.binary_mask_dilation(hole))
    ref, _ = nb.diagnostics.C_eig_2D(nb.integration.flowmap_grid_2D(f, 0.0, 5.0, x, y, p), x[1], y[1])
    keep = ~U.binary_mask_dilation(hole)
    assert np.array_equal(vals[keep], ref[keep]) and not vals[~keep].any()


# ------------------------------------------------------------------ edge cases

def test_edge_cases_of_the_tensor_and_ridge_entries(nb, oracle):
    D, E, I = nb.diagnostics, nb.extraction, nb.integration
    # empty grids
    vals, vecs = D.C_eig_2D(np.zeros((0, 5, 2)), 0.1, 0.1)
    assert vals.shape == (0, 5, 2) and vecs.shape == (0, 5, 2, 2)
    assert D.C_tensor_2D(np.zeros((4, 0, 5, 2)), 0.1, 0.1).shape == (4, 0, 3)
    assert D.ftle_from_eig(np.zeros((0, 3)), 1.0).shape == (0, 3)
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    assert I.flowmap_aux_grid_2D(f, 0.0, 1.0, np.zeros(0), np.linspace(0, 1, 4), p).shape == (0, 4, 5, 2)
    # grids smaller than the stencils: everything is border -> zeros / no ridge points
    rng = np.random.default_rng(4)
    for shape in ((2, 2), (3, 4), (4, 4)):
        fm = rng.normal(size=shape + (2,))
        vals, vecs = D.C_eig_2D(fm, 0.1, 0.2)
        vo, eo = oracle.C_eig_2D(fm, 0.1, 0.2)
        assert np.array_equal(vals, vo) and np.array_equal(vecs, eo)
        fa = rng.normal(size=shape + (5, 2))
        assert not D.C_tensor_2D(fa, 0.1, 0.2).any()
        assert not D.C_eig_aux_2D(fa, 0.1, 0.2)[0].any()
        x, y = np.linspace(0, 1, shape[0]), np.linspace(0, 1, shape[1])
        r = E.ftle_ridge_pts(rng.random(shape) + 1, rng.normal(size=shape + (2,)), x, y)
        assert r.shape == (0, 2)
        assert E.ftle_ridges(rng.random(shape) + 1, rng.normal(size=shape + (2,)), x, y) == []
    # a field without ridges (monotone ramp), and one where every interior pixel is a ridge point
    x, y = np.linspace(0, 1, 40), np.linspace(0, 1, 30)
    X, Y = np.meshgrid(x, y, indexing="ij")
    ev = np.zeros(X.shape + (2,))
    ev[..., 0] = 1.0
    assert E.ftle_ridge_pts(1 + X + Y, ev, x, y).shape == (0, 2)
    assert E.ftle_ridge_pts(2 - (X - 0.5) ** 2, ev, x, y, percentile=100).shape == (0, 2)   # f > max(f) never
    bump = 2 - 1e-3 * (X - X.round(1)) ** 2 - (X - 0.5) ** 2        # concave in x everywhere
    r = E.ftle_ridge_pts(bump, ev, x, y)
    ro = oracle.ftle_ridge_pts(bump, ev, x, y)
    assert np.array_equal(r, ro)
    # more ridge points than the wrapper's first capacity guess (max(4096, pixels / 32)): second call
    xs, ys = np.linspace(0, 1, 300), np.linspace(0, 1, 300)
    Xs, Ys = np.meshgrid(xs, ys, indexing="ij")
    wavy = 2 + np.cos(2 * np.pi * Xs / (3 * (xs[1] - xs[0])))      # a ridge every third column
    evs = np.zeros(Xs.shape + (2,))
    evs[..., 0] = 1.0
    r = E.ftle_ridge_pts(wavy, evs, xs, ys)
    ro = oracle.ftle_ridge_pts(wavy, evs, xs, ys)
    assert len(ro) > max(4096, 300 * 300 // 32) and np.array_equal(r, ro)
    rr = E.ftle_ridges(wavy, evs, xs, ys)
    rro = oracle.ftle_ridges(wavy, evs, xs, ys)
    assert [len(a) for a in rr] == [len(a) for a in rro] and np.array_equal(np.concatenate(rr), np.concatenate(rro))
    # argument errors
    with pytest.raises(ValueError):
        D.C_eig_2D(np.zeros((4, 4, 3)), 0.1, 0.1)
    with pytest.raises(ValueError):
        D.C_eig_aux_2D(np.zeros((6, 6, 4, 2)), 0.1, 0.1, eig_main=True)
    with pytest.raises(ValueError):
        E.ftle_ridge_pts(np.zeros((6, 6)), np.zeros((6, 5, 2)), np.arange(6.0), np.arange(6.0))
    with pytest.raises(ValueError):
        I.flowmap_composition(np.zeros((3, 5, 5, 2)), ((0, 1, 5), (0, 1, 6)), 3)
    with pytest.raises(NotImplementedError):
        I.flowmap_aux_grid_2D(f, 0.0, 1.0, x, y, p, method="lsoda")


# ------------------------------------------------------------------ time series in one launch

def test_flowmap_and_ftle_series_equal_the_per_frame_calls(nb):
    import torch
    I, D = nb.integration, nb.diagnostics
    x, y = np.linspace(0, 2, 67), np.linspace(0, 1, 35)
    dx, dy = x[1] - x[0], y[1] - y[0]
    t0s = np.array([0.0, 0.5, 3.25, 7.0, -2.0])
    rng = np.random.default_rng(8)
    mask = rng.random((67, 35)) < 0.1
    for direction, T in ((1.0, 6.0), (-1.0, -6.0)):
        f, p, _ = nb.flows.get_predefined_flow("double_gyre", int_direction=direction)
        info = {}
        fms = I.flowmap_grid_2D_series(f, t0s, T, x, y, p, mask=mask, info=info)
        assert fms.shape == (5, 67, 35, 2) and info["status"].shape == (5, 67, 35)
        fts = D.ftle_grid_2D_series(fms, T, dx, dy, mask=mask)
        tot = np.zeros(3, np.int64)
        for k, t0 in enumerate(t0s):
            one = {}
            fm = I.flowmap_grid_2D(f, t0, T, x, y, p, mask=mask, info=one)
            assert np.array_equal(fms[k], fm) and np.array_equal(info["steps"][k], one["steps"])
            assert np.array_equal(fts[k], D.ftle_grid_2D(fm, T, dx, dy, mask=mask))
            tot += one["stats"]
        assert np.array_equal(info["stats"], tot)
    # device-resident, other flows, writing into a preallocated tensor
    fb, pb, dom = nb.flows.get_predefined_flow("bickley_jet")
    xb = torch.linspace(dom[0][0], dom[0][1], 50, dtype=torch.float64, device="cuda")
    yb = torch.linspace(-3, 3, 30, dtype=torch.float64, device="cuda")
    out = torch.empty((3, 50, 30, 2), dtype=torch.float64, device="cuda")
    res = I.flowmap_grid_2D_series(fb, [0.0, 1.0, 2.0], 3.0, xb, yb, pb, out=out)
    assert res is out
    for k in range(3):
        assert torch.equal(out[k], I.flowmap_grid_2D(fb, float(k), 3.0, xb, yb, pb))
    assert I.flowmap_grid_2D_series(fb, np.zeros(0), 3.0, xb, yb, pb).shape == (0, 50, 30, 2)
    with pytest.raises(ValueError):
        I.flowmap_grid_2D_series(fb, [0.0], 3.0, xb, yb, pb, out=torch.empty((2, 50, 30, 2), dtype=torch.float64, device="cuda"))


def test_flowmap_composition_series_equals_initial_plus_steps(nb):
    import torch
    I = nb.integration
    nx, ny, n = 81, 41, 7
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    grid = ((x[0], x[-1], nx), (y[0], y[-1], ny))
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    t0, T, h = 0.5, 6.0, 1.5
    series = I.flowmap_composition_series(f, t0, T, h, n, x, y, grid, p)
    assert series.shape == (n, nx, ny, 2)
    fm0, fms, nT = I.flowmap_composition_initial(f, t0, T, h, x, y, grid, p)
    assert nT == 4 and np.array_equal(series[0], fm0)
    for k in range(1, n):
        fmk, fms = I.flowmap_composition_step(fms, f, t0 + T + (k - 1) * h, h, nT, x, y, grid, p)
        assert np.array_equal(series[k], fmk)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    sd = I.flowmap_composition_series(f, t0, T, h, n, xd, yd, grid, p)
    assert sd.is_cuda and np.array_equal(sd.cpu().numpy(), series)


# ------------------------------------------------------------------ ordered ridges (section 8(f)1)

@pytest.mark.parametrize("tag", ["ref", "dg_a", "dg_b", "dg_c", "rnd_a", "rnd_b", "rnd_c"])
def test_ftle_ordered_ridges_match_the_real_reference(nb, tag):
    """ftle_ordered_ridges end to end (per-pixel ridge test on the GPU, linking + end-point matching in
    the library's host code) against frozen outputs of the real reference functions
    (tests/golden/make_ordered_ridges_golden.py) and, for `ref`, the reference's own pickled golden
    (tests/test_extraction.py:23-29)."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ordered_ridges_golden.npz"))
    dist_tol, ang, mrp, thr, pct, c, h = G[tag + "_args"]
    f, ev, x, y = (G[f"{tag}_{k}"] for k in ("f", "ev", "x", "y"))
    lk, rl, ep, tv = nb.extraction._linked_ridge_pts(f, ev, x, y, sdd_thresh=thr, percentile=int(pct), c=c)
    assert np.array_equal(rl, G[tag + "_ridge_len"]) and np.array_equal(lk, G[tag + "_linked"])
    assert np.array_equal(ep, G[tag + "_endpoints"]) and np.allclose(tv, G[tag + "_tanvecs"], rtol=0, atol=1e-15)
    ridges = nb.extraction.ftle_ordered_ridges(f, ev, x, y, dist_tol, ep_tan_ang=ang, min_ridge_pts=int(mrp),
                                               sdd_thresh=thr, percentile=int(pct), c=c)
    assert [len(r) for r in ridges] == list(G[tag + "_ordered_len"])
    assert np.array_equal(np.concatenate(ridges) if ridges else np.zeros((0, 2)), G[tag + "_ordered_cat"])
    if tag == "ref":
        assert [len(r) for r in ridges] == list(G["ref_pkl_ordered_len"])
        assert np.allclose(np.concatenate(ridges), G["ref_pkl_ordered_cat"])


def test_dg_ftle_ridges_example_runs_as_numbacs(nb, oracle):
    """examples/ftle/plot_dg_ftle_ridges.py:13-72 of the reference, its compute lines verbatim
    (imports through the `numbacs` alias, matplotlib left out), on the GPU path; the ridges are
    compared with the same pipeline fed by the CPU oracle's flow map."""
    nb.install_as_numbacs()
    from math import copysign
    from numbacs.flows import get_predefined_flow
    from numbacs.integration import flowmap_grid_2D
    from numbacs.diagnostics import ftle_from_eig, C_eig_2D
    from numbacs.extraction import ftle_ordered_ridges
    t0 = 0.0
    T = -10.0
    int_direction = copysign(1, T)
    funcptr, params, domain = get_predefined_flow("double_gyre", int_direction=int_direction)
    nx, ny = 401, 201
    x = np.linspace(domain[0][0], domain[0][1], nx)
    y = np.linspace(domain[1][0], domain[1][1], ny)
    dx = x[1] - x[0]
    dy = y[1] - y[0]
    flowmap = flowmap_grid_2D(funcptr, t0, T, x, y, params)
    eigvals, eigvecs = C_eig_2D(flowmap, dx, dy)
    eigval_max = eigvals[:, :, 1]
    eigvec_max = eigvecs[:, :, :, 1]
    ftle = ftle_from_eig(eigval_max, T)
    percentile = 0
    sdd_thresh = 10.0
    dist_tol = 5e-2
    ridge_curves = ftle_ordered_ridges(
        ftle, eigvec_max, x, y, dist_tol, percentile=percentile, sdd_thresh=sdd_thresh
    )
    assert len(ridge_curves) > 10 and all(r.ndim == 2 and r.shape[1] == 2 and len(r) >= 5 for r in ridge_curves)
    assert max(len(r) for r in ridge_curves) > 300           # the main ridge is one long ordered curve
    # the same tail on the oracle's flow map
    fo, po, _ = oracle.get_predefined_flow("double_gyre", int_direction=int_direction)
    fmo = oracle.flowmap_grid_2D(fo, t0, T, x, y, po)
    vo, eo = C_eig_2D(fmo, dx, dy)
    ref_curves = ftle_ordered_ridges(ftle_from_eig(vo[:, :, 1], T), eo[:, :, :, 1], x, y, dist_tol,
                                     percentile=percentile, sdd_thresh=sdd_thresh)
    n_gpu, n_ref = sum(len(r) for r in ridge_curves), sum(len(r) for r in ref_curves)
    print("ordered ridges gpu/oracle:", len(ridge_curves), len(ref_curves), "points", n_gpu, n_ref)
    assert abs(n_gpu - n_ref) <= 0.01 * n_ref and abs(len(ridge_curves) - len(ref_curves)) <= 2
    if [len(r) for r in ridge_curves] == [len(r) for r in ref_curves]:
        assert np.abs(np.concatenate(ridge_curves) - np.concatenate(ref_curves)).max() <= 1e-6
