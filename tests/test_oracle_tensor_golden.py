"""CPU tests: the oracle's aux-grid flow map, Cauchy-Green tensor / eigen-pairs and FTLE ridge
points against the reference's goldens and against outputs of the real reference code frozen in
tests/golden/reference_golden.npz.

Mirrors /root/reference/tests/test_integration.py:66-75, test_diagnostics.py:18-47, 99-150,
test_extraction.py:7-13."""
import numpy as np


def apply_mask(arr, mask):
    out = arr.copy()
    out[mask] = 0.0
    return out


def reconstruct_matrix(evals, evecs):
    """V diag(w) V^T per pixel (test_diagnostics.py:18-47): independent of eigenvector signs."""
    return np.einsum("...ik,...k,...jk->...ij", evecs, evals, evecs)


def test_flowmap_aux_grid_2D_golden(oracle, golden, coords_dg, mask_dg):
    x, y = coords_dg
    f, p, _ = oracle.get_predefined_flow("double_gyre")
    fa = oracle.flowmap_aux_grid_2D(f, 0.0, 8.0, x, y, p)
    assert fa.shape == (21, 11, 5, 2)
    assert np.array_equal(fa.astype(np.float32), golden["ref_fm_aux"])
    # centre point == plain flow map, bit for bit (same particle, same solver)
    assert np.array_equal(fa[:, :, 4, :], oracle.flowmap_grid_2D(f, 0.0, 8.0, x, y, p))
    # edge cells only carry the centre point (integration.py:320-343)
    assert not fa[0, :, :4].any() and not fa[:, 0, :4].any() and not fa[-1, :, :4].any()
    fa_m = oracle.flowmap_aux_grid_2D(f, 0.0, 8.0, x, y, p, mask=mask_dg)
    assert np.array_equal(fa_m.astype(np.float32), apply_mask(golden["ref_fm_aux"], mask_dg))


def test_flowmap_aux_grid_2D_variants(oracle, coords_dg):
    x, y = coords_dg
    f, p, _ = oracle.get_predefined_flow("double_gyre")
    full = oracle.flowmap_aux_grid_2D(f, 0.0, 8.0, x, y, p)
    inner = oracle.flowmap_aux_grid_2D(f, 0.0, 8.0, x, y, p, compute_edge=False)
    assert np.array_equal(inner[1:-1, 1:-1], full[1:-1, 1:-1])
    assert not inner[0].any() and not inner[-1].any() and not inner[:, 0].any() and not inner[:, -1].any()
    four = oracle.flowmap_aux_grid_2D(f, 0.0, 8.0, x, y, p, eig_main=False)
    assert four.shape == (21, 11, 4, 2)
    assert np.array_equal(four[1:-1, 1:-1], full[1:-1, 1:-1, :4])
    assert four[0, :, 2:, 1].all() and four[:, -1, :, 1].all()   # edges integrated too when compute_edge
    four_in = oracle.flowmap_aux_grid_2D(f, 0.0, 8.0, x, y, p, eig_main=False, compute_edge=False)
    assert np.array_equal(four_in[1:-1, 1:-1], four[1:-1, 1:-1]) and not four_in[0].any()


def test_C_tensor_2D_golden(oracle, golden, coords_dg, mask_dg):
    x, y = coords_dg
    fa = golden["ref_fm_aux"].astype(np.float64)
    C = oracle.C_tensor_2D(fa, x[1], y[1])
    assert np.allclose(C.astype(np.float32), golden["ref_C"])
    Cm = oracle.C_tensor_2D(fa, x[1], y[1], mask=mask_dg)
    assert np.allclose(Cm.astype(np.float32), apply_mask(golden["ref_C"], mask_dg))


def test_C_eig_2D_golden(oracle, golden, coords_dg, mask_dg):
    x, y = coords_dg
    fm = golden["ref_fm"].astype(np.float64)
    Cexp = reconstruct_matrix(golden["ref_Cevals"], golden["ref_Cevecs"])
    vals, vecs = oracle.C_eig_2D(fm, x[1], y[1])
    assert np.allclose(reconstruct_matrix(vals.astype(np.float32), vecs.astype(np.float32)), Cexp)
    vals, vecs = oracle.C_eig_2D(fm, x[1], y[1], mask_dg)
    assert np.allclose(reconstruct_matrix(vals.astype(np.float32), vecs.astype(np.float32)),
                       apply_mask(Cexp, mask_dg))


def test_C_eig_aux_2D_golden(oracle, golden, coords_dg, mask_dg):
    x, y = coords_dg
    fa = golden["ref_fm_aux"].astype(np.float64)
    Cexp = reconstruct_matrix(golden["ref_Cevals_aux"], golden["ref_Cevecs_aux"])
    vals, vecs = oracle.C_eig_aux_2D(fa, x[1], y[1])
    assert np.allclose(reconstruct_matrix(vals.astype(np.float32), vecs.astype(np.float32)), Cexp)
    vals, vecs = oracle.C_eig_aux_2D(fa, x[1], y[1], mask=mask_dg)
    assert np.allclose(reconstruct_matrix(vals.astype(np.float32), vecs.astype(np.float32)),
                       apply_mask(Cexp, mask_dg))


def test_tensor_functions_match_real_reference_bitwise(oracle, golden):
    """Outputs of the real numbacs.diagnostics code (LAPACK eigh through numba) on seeded float64
    inputs: eigenvalues, eigenvector components AND signs are reproduced bit for bit."""
    dx, dy, h = golden["ceig_args"]
    fm, fa, mask = golden["ceig_in"], golden["caux_in"], golden["ceig_mask"]
    for tag, m in (("", None), ("_masked", mask)):
        vals, vecs = oracle.C_eig_2D(fm, dx, dy, m)
        assert np.array_equal(vals, golden["ceig_vals" + tag])
        assert np.array_equal(vecs, golden["ceig_vecs" + tag])
        assert np.array_equal(oracle.C_tensor_2D(fa, dx, dy, h, m), golden["ctensor" + tag])
        vals, vecs = oracle.C_eig_aux_2D(fa, dx, dy, h, True, m)
        assert np.array_equal(vals, golden["caux_vals_main" + tag])
        assert np.array_equal(vecs, golden["caux_vecs_main" + tag])
        vals, vecs = oracle.C_eig_aux_2D(fa[:, :, :4], dx, dy, h, False, m)
        assert np.array_equal(vals, golden["caux_vals" + tag])
        assert np.array_equal(vecs, golden["caux_vecs" + tag])
    assert np.array_equal(oracle.ftle_from_eig(golden["ceig_vals"][:, :, 1], -3.0),
                          golden["ftle_from_eig_out"])


def test_eigh2_against_numpy_lapack(oracle):
    """The dlaev2 restatement against numpy's LAPACK on random and degenerate symmetric 2x2s."""
    rng = np.random.default_rng(7)
    mats = [rng.normal(size=(2, 2)) * 10 ** rng.uniform(-3, 3) for _ in range(2000)]
    mats = [F.T @ F for F in mats]
    mats += [np.zeros((2, 2)), np.diag([3.0, 1.0]), np.diag([1.0, 3.0]), np.diag([2.0, 2.0]),
             np.array([[3.0, 1e-17], [1e-17, 1.0]]), np.array([[1.0, 1e-8], [1e-8, 3.0]]),
             np.array([[3.0, 4e-16], [4e-16, 1.0]]), np.array([[2.0, 1e-9], [1e-9, 2.0]])]
    for C in mats:
        w, v = np.linalg.eigh(C)
        wo, vo = oracle.eigh2(C[0, 0], C[0, 1], C[1, 1])
        assert np.array_equal(w, wo) and np.array_equal(v, vo), (C, w, wo, v, vo)


def test_ftle_ridge_pts_golden(oracle, golden, coords_dg):
    x, y = coords_dg
    r = oracle.ftle_ridge_pts(golden["ref_ftle"], golden["ref_Cevecs"][:, :, :, 1], x, y)
    assert r.shape == golden["ref_ridge_pts"].shape
    assert np.allclose(r, golden["ref_ridge_pts"])


def test_ftle_ridge_pts_match_real_reference_bitwise(oracle, golden):
    f, ev, x, y = golden["ridge_f"], golden["ridge_ev"], golden["ridge_x"], golden["ridge_y"]
    for tag, (thr, pct) in zip("abc", golden["ridge_args"]):
        r = oracle.ftle_ridge_pts(f, ev, x, y, thr, int(pct))
        assert np.array_equal(r, golden["ridge_pts_" + tag])
        rp, rv, sdd, _ = oracle._ftle_ridge_pts_connect(f, ev, x, y, thr, int(pct))
        assert np.array_equal(rp, golden["ridge_conn_pts_" + tag])
        assert np.array_equal(rv, golden["ridge_conn_vec_" + tag])
        assert np.array_equal(sdd, golden["ridge_conn_sdd_" + tag])


def _check_ridges(ridges, cat, lens, exact):
    assert [len(r) for r in ridges] == list(lens)
    got = np.concatenate(ridges)
    assert np.array_equal(got, cat) if exact else np.allclose(got, cat)


def test_ftle_ridges_golden(oracle, golden, coords_dg):
    """tests/test_extraction.py:15-21 of the reference (ridges.pkl)."""
    x, y = coords_dg
    r = oracle.ftle_ridges(golden["ref_ftle"], golden["ref_Cevecs"][:, :, :, 1], x, y)
    _check_ridges(r, golden["ref_ridges_cat"], golden["ref_ridges_len"], exact=False)


def test_ftle_ridges_match_real_reference(oracle, golden):
    f, ev, x, y = golden["ridge_f"], golden["ridge_ev"], golden["ridge_x"], golden["ridge_y"]
    for tag, (thr, pct, mrp) in zip("ab", golden["ridges_args"]):
        r = oracle.ftle_ridges(f, ev, x, y, thr, int(pct), int(mrp))
        _check_ridges(r, golden["ridges_cat_" + tag], golden["ridges_len_" + tag], exact=True)


def check_composition_initial(fm0, fms, nT, golden):
    """fm_ci.npy / fms_ci.npy.  Wall particles are compared loosely: they sit at the wall
    coordinate +- a few 1e-16 (rounding noise of the wall-normal velocity, which depends on the
    sin/cos implementation: the reference's golden machine, glibc here, the GPU's own), and when
    an interpolated position lands one ulp outside the grid the CONSTANT extrapolation returns 0
    -- fm_ci.npy itself holds such zeros along x = 2.  On the walls an entry may therefore be
    either the golden value or (0, 0); the interior must match."""
    assert nT == 8
    assert np.array_equal(fms.astype(np.float32), golden["ref_fms_ci"])
    ref = golden["ref_fm_ci"]
    got = fm0.astype(np.float32)
    assert np.allclose(got[1:-1, 1:-1], ref[1:-1, 1:-1])
    wall = np.ones(ref.shape[:2], bool)
    wall[1:-1, 1:-1] = False
    ok = np.isclose(got, ref).all(-1) | (got == 0).all(-1) | (ref == 0).all(-1)
    assert ok[wall].all()


def test_flowmap_composition_golden(oracle, golden, coords_dg):
    """tests/test_integration.py:92-114 of the reference (fm_ci / fms_ci / fm_cs / fms_cs.npy)."""
    x, y = coords_dg
    grid = ((x[0], x[-1], 21), (y[0], y[-1], 11))
    f, p, _ = oracle.get_predefined_flow("double_gyre")
    fm0, fms, nT = oracle.flowmap_composition_initial(f, 0.0, 8.0, 1.0, x, y, grid, p)
    check_composition_initial(fm0, fms, nT, golden)
    fmk, fms2 = oracle.flowmap_composition_step(golden["ref_fms_ci"].astype(np.float64), f, 8.0, 1.0,
                                                8, x, y, grid, p)
    assert np.allclose(fmk.astype(np.float32), golden["ref_fm_cs"])
    assert np.allclose(fms2.astype(np.float32), golden["ref_fms_cs"])
    # the composed map approximates the directly integrated one (it is an interpolation scheme)
    direct = oracle.flowmap_grid_2D(f, 0.0, 8.0, x, y, p)
    assert np.abs(fm0 - direct).max() < 0.5 and np.median(np.abs(fm0 - direct)) < 0.05
    # points that leave the grid get 0 (CONSTANT extrapolation), per component
    fms3 = fms.copy()
    fms3[0, 3, 4] = (2.5, 0.5)
    fms3[0, 5, 6] = (1.0, -1e-9)
    out = oracle.flowmap_composition(fms3, grid, 8)
    assert not out[3, 4].any() and not out[5, 6].any()
    assert np.array_equal(np.delete(out.reshape(-1, 2), [3 * 11 + 4, 5 * 11 + 6], axis=0),
                          np.delete(fm0.reshape(-1, 2), [3 * 11 + 4, 5 * 11 + 6], axis=0))


def test_binary_mask_dilation_matches_real_reference(oracle, golden):
    assert np.array_equal(oracle.binary_mask_dilation(golden["dil_in"]), golden["dil_out4"])
    assert np.array_equal(oracle.binary_mask_dilation(golden["dil_in"], corners=True), golden["dil_out8"])
