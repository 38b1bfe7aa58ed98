#!/usr/bin/env python
"""Generate tests/golden/reference_golden.npz from the reference tree (run in the build container).

    python tests/golden/make_golden.py [/root/reference]

The reference cannot travel to the GPU box, so everything the parity tests need from it is
frozen here as small arrays:

 (1) the reference's own float32 golden vectors for the hot path
     (/root/reference/tests/testing_data/{fm,fm_n,ftle,lavd,vort}.npy);
 (2) the literal known-answer tables in /root/reference/tests/test_flows.py (spline prefilter
     coefficients, cubic/linear evaluations, double-gyre / bickley / abc velocity tables) and
     the scalar KATs in tests/test_utils.py -- pulled out with `ast`, never retyped;
 (3) outputs of the REAL reference code that imports here (numbacs.diagnostics.ftle_grid_2D,
     numbacs.utils.composite_simpsons / eigvalsh_max_2D run from /root/reference/src) on seeded
     float64 inputs, so FTLE parity is checked beyond float32 resolution;
 (4) likewise numbacs.diagnostics.{C_eig_2D, C_tensor_2D, C_eig_aux_2D, ftle_from_eig} and the
     ridge-point stencil of numbacs/extraction/ridges.py.

Nothing in this script is used at test time; only the .npz is.
"""
import ast
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def literal_arrays(path, func_name, var_names):
    """Evaluate `var = np.array(...)` / tuple-of-np.array assignments inside one test function."""
    tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == func_name:
            for st in ast.walk(node):
                if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) \
                        and st.targets[0].id in var_names:
                    code = compile(ast.Expression(st.value), path, "eval")
                    out[st.targets[0].id] = eval(code, {"np": np, "pi": np.pi})
    missing = set(var_names) - set(out)
    assert not missing, (func_name, missing)
    return out


def main():
    G = {}
    td = os.path.join(REF, "tests", "testing_data")
    for name in ("fm", "fm_n", "ftle", "lavd", "vort", "fm_aux", "C", "Cevals", "Cevecs",
                 "Cevals_aux", "Cevecs_aux", "ridge_pts", "fm_ci", "fms_ci", "fm_cs", "fms_cs"):
        G["ref_" + name] = np.load(os.path.join(td, name + ".npy"))

    tf = os.path.join(REF, "tests", "test_flows.py")
    a = literal_arrays(tf, "test_get_interp_arrays_2D", ["Cu_expected", "Cv_expected"])
    G["spline_Cu"], G["spline_Cv"] = a["Cu_expected"], a["Cv_expected"]
    a = literal_arrays(tf, "test_get_callable_2D", ["veli_expected"])
    G["spline_eval_u"], G["spline_eval_v"] = a["veli_expected"]
    a = literal_arrays(tf, "test_get_callable_linear_2D", ["veli_expected"])
    G["linear_eval_u"], G["linear_eval_v"] = a["veli_expected"]
    a = literal_arrays(tf, "test_get_predefined_callable_dg", ["veli_expected"])
    G["vel_dg_u"], G["vel_dg_v"] = a["veli_expected"]
    a = literal_arrays(tf, "test_get_predefined_callable_bickley", ["veli_expected"])
    G["vel_bickley_u"], G["vel_bickley_v"] = a["veli_expected"]
    a = literal_arrays(tf, "test_get_predefined_callable_abc", ["veli_expected"])
    G["vel_abc_u"], G["vel_abc_v"], G["vel_abc_w"] = a["veli_expected"]

    # (3) real reference code on seeded inputs
    sys.path.insert(0, os.path.join(REF, "src"))
    from numbacs.diagnostics import ftle_grid_2D
    from numbacs.utils import composite_simpsons, eigvalsh_max_2D

    rng = np.random.default_rng(20261017)
    fm = rng.normal(size=(37, 29, 2)) * 2.0
    mask = rng.random((37, 29)) < 0.2
    G["ftle_in"] = fm
    G["ftle_mask"] = mask
    G["ftle_args"] = np.array([-7.5, 0.05, 0.0125])  # T, dx, dy
    G["ftle_out"] = ftle_grid_2D(fm, -7.5, 0.05, 0.0125)
    G["ftle_out_masked"] = ftle_grid_2D(fm, -7.5, 0.05, 0.0125, mask=mask)
    # near-identity map: exercises the max_eig <= 1 branch (diagnostics.py:62)
    X, Y = np.meshgrid(np.linspace(0, 1, 16), np.linspace(0, 1, 12), indexing="ij")
    fm2 = np.stack([0.9 * X, 0.8 * Y], axis=-1)
    G["ftle_in_contract"] = fm2
    G["ftle_out_contract"] = ftle_grid_2D(fm2, 3.0, X[1, 0] - X[0, 0], Y[0, 1] - Y[0, 0])

    f_even = rng.normal(size=11)
    f_odd = rng.normal(size=12)
    G["simpson_in_even"], G["simpson_in_odd"] = f_even, f_odd
    G["simpson_out"] = np.array([composite_simpsons(f_even, 0.3), composite_simpsons(f_odd, 0.3)])
    A = np.array([[2.5, -1.25], [-1.25, 0.75]])
    G["eig_in"] = A
    G["eig_out"] = np.array([eigvalsh_max_2D(A)])

    # (4) the Cauchy-Green tensor / eigen-pair functions (real numbacs.diagnostics code) and the
    # ridge-point stencil (real numbacs/extraction/ridges.py, loaded without the package
    # __init__, which needs the missing `interpolation` package) on seeded float64 inputs
    import importlib.util
    import types
    from numbacs.diagnostics import C_eig_2D, C_eig_aux_2D, C_tensor_2D, ftle_from_eig

    fm3 = rng.normal(size=(23, 19, 2)) * 3.0
    fa3 = rng.normal(size=(23, 19, 5, 2)) * 1e-4
    mask3 = rng.random((23, 19)) < 0.2
    G["ceig_in"], G["caux_in"], G["ceig_mask"] = fm3, fa3, mask3
    G["ceig_args"] = np.array([0.1, 0.07, 1e-5])  # dx, dy, h
    for tag, m in (("", None), ("_masked", mask3)):
        G["ceig_vals" + tag], G["ceig_vecs" + tag] = C_eig_2D(fm3, 0.1, 0.07, m)
        G["ctensor" + tag] = C_tensor_2D(fa3, 0.1, 0.07, 1e-5, m)
        G["caux_vals_main" + tag], G["caux_vecs_main" + tag] = C_eig_aux_2D(fa3, 0.1, 0.07, 1e-5, True, m)
        G["caux_vals" + tag], G["caux_vecs" + tag] = C_eig_aux_2D(fa3[:, :, :4], 0.1, 0.07, 1e-5, False, m)
    G["ftle_from_eig_out"] = ftle_from_eig(G["ceig_vals"][:, :, 1], -3.0)

    pkg = types.ModuleType("numbacs.extraction")
    pkg.__path__ = [os.path.join(REF, "src", "numbacs", "extraction")]
    sys.modules["numbacs.extraction"] = pkg
    spec = importlib.util.spec_from_file_location(
        "numbacs.extraction.ridges", os.path.join(pkg.__path__[0], "ridges.py"))
    ridges = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = ridges
    spec.loader.exec_module(ridges)
    xr, yr = np.linspace(0, 2, 60), np.linspace(0, 1, 45)
    X, Y = np.meshgrid(xr, yr, indexing="ij")
    f = np.sin(3 * X + Y * Y) * np.cos(2 * Y) + 0.03 * rng.normal(size=X.shape) + 0.5
    th = rng.uniform(0, 2 * np.pi, size=X.shape)
    ev = np.stack([np.cos(th), np.sin(th)], -1)
    G["ridge_f"], G["ridge_ev"], G["ridge_x"], G["ridge_y"] = f, ev, xr, yr
    for tag, (thr, pct) in (("a", (0.0, 0)), ("b", (10.0, 0)), ("c", (0.0, 60))):
        G["ridge_pts_" + tag] = ridges.ftle_ridge_pts(f, ev, xr, yr, thr, pct)
        rp, rv, sdd, hmin = ridges._ftle_ridge_pts_connect(f, ev, xr, yr, thr, pct)
        G["ridge_conn_pts_" + tag], G["ridge_conn_vec_" + tag], G["ridge_conn_sdd_" + tag] = rp, rv, sdd
    G["ridge_args"] = np.array([[0.0, 0], [10.0, 0], [0.0, 60]])
    # connected ridges: the reference's pickled golden (tests/test_extraction.py:15-21) and the
    # real ftle_ridges on the seeded field, stored as one stacked array + the ridge lengths
    import pickle
    with open(os.path.join(td, "ridges.pkl"), "rb") as fh:
        rd = pickle.load(fh)
    G["ref_ridges_cat"], G["ref_ridges_len"] = np.concatenate(rd), np.array([len(r) for r in rd])
    for tag, (thr, pct, mrp) in (("a", (0.0, 0, 3)), ("b", (10.0, 60, 1))):
        rr = ridges.ftle_ridges(f, ev, xr, yr, thr, pct, mrp)
        G["ridges_cat_" + tag] = np.concatenate(rr)
        G["ridges_len_" + tag] = np.array([len(r) for r in rr])
    G["ridges_args"] = np.array([[0.0, 0, 3], [10.0, 60, 1]])

    # (5) mask dilation: the real numbacs.utils.binary_mask_dilation on a seeded mask
    from numbacs.utils import binary_mask_dilation
    mk = rng.random((37, 29)) < 0.08
    mk[0, 0] = mk[-1, -1] = mk[0, 13] = mk[20, -1] = True
    G["dil_in"] = mk
    G["dil_out4"] = binary_mask_dilation(mk)
    G["dil_out8"] = binary_mask_dilation(mk, corners=True)

    out = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out), "bytes;", len(G), "arrays")


if __name__ == "__main__":
    main()
