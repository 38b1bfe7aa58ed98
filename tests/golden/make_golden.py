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
     float64 inputs, so FTLE parity is checked beyond float32 resolution.

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
    for name in ("fm", "fm_n", "ftle", "lavd", "vort"):
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

    out = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, os.path.getsize(out), "bytes;", len(G), "arrays")


if __name__ == "__main__":
    main()
