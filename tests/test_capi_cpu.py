"""CPU tests of the boundary: libb200cs.so builds, loads, exports every symbol include/b200cs.h
declares, the ctypes prototypes cover them, and the product fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "b200cs.h")).read()
    return sorted(set(re.findall(r"B200CS_API\s+[\w\s\*]+?\b(b200cs_\w+)\s*\(", hdr)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for needed in ("b200cs_flowmap_grid_2d", "b200cs_flowmap_pts", "b200cs_ftle_grid_2d",
                   "b200cs_lavd_grid_2d", "b200cs_flow_create_analytic",
                   "b200cs_flow_create_spline", "b200cs_prefilter_3d", "b200cs_last_error"):
        assert needed in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"libb200cs.so does not export {name}"


def test_ctypes_prototypes_cover_header(lib):
    from numbacs_b200 import _lib
    declared = set(declared_symbols()) - {"b200cs_last_error"}
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.b200cs_version() == 100


def test_flow_registry_without_gpu(lib):
    """Handles and metadata are host-side state: usable without a device."""
    from numbacs_b200.flows import get_predefined_flow, release_flow
    import ctypes as C
    f, p, dom = get_predefined_flow("double_gyre", int_direction=-1.0)
    assert isinstance(f, int) and p[0] == -1.0 and dom == ((0.0, 2.0), (0.0, 1.0))
    assert np.allclose(p, [-1.0, 0.1, 0.25, 0.0, 0.2 * np.pi, 0.0])
    kind, ndim, npar = C.c_int(), C.c_int(), C.c_int()
    assert lib.b200cs_flow_info(f, C.byref(kind), C.byref(ndim), C.byref(npar)) == 0
    assert (kind.value, ndim.value, npar.value) == (0, 2, 6)
    fb, pb, domb = get_predefined_flow("bickley_jet", int_direction=-1.0)
    assert pb[0] == 1.0 and len(pb) == 12  # the reference forces +1 (flows.py:1219)
    fa, pa, doma, desc = get_predefined_flow("abc", parameter_description=True)
    assert len(pa) == 5 and len(doma) == 3 and desc.startswith("p[0] = int_direction")
    assert isinstance(get_predefined_flow("abc", return_default_params=False,
                                          return_domain=False), int)
    release_flow(f)
    assert lib.b200cs_flow_info(f, None, None, None) == -3
    with pytest.raises(ValueError):
        get_predefined_flow("nope")


def _has_gpu():
    try:
        from numbacs_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from numbacs_b200.flows import get_predefined_flow
    from numbacs_b200.integration import flowmap_grid_2D
    from numbacs_b200.diagnostics import ftle_grid_2D
    f, p, _ = get_predefined_flow("double_gyre")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        flowmap_grid_2D(f, 0.0, 8.0, np.linspace(0, 2, 5), np.linspace(0, 1, 5), p)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ftle_grid_2D(np.zeros((5, 5, 2)), 8.0, 0.1, 0.1)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_for_the_tensor_and_ridge_entries(lib):
    """Every compute entry added for the section-8(f) rows fails loudly without a GPU."""
    import numbacs_b200 as nb
    f, p, _ = nb.flows.get_predefined_flow("double_gyre")
    x, y = np.linspace(0, 2, 7), np.linspace(0, 1, 6)
    fm, fa = np.zeros((7, 6, 2)), np.zeros((7, 6, 5, 2))
    grid = ((0.0, 2.0, 7), (0.0, 1.0, 6))
    calls = [
        lambda: nb.integration.flowmap_aux_grid_2D(f, 0.0, 1.0, x, y, p),
        lambda: nb.integration.flowmap_grid_ND(f, 0.0, 1.0, np.zeros(8), 2, p),
        lambda: nb.integration.flowmap_composition(np.zeros((3, 7, 6, 2)), grid, 3),
        lambda: nb.integration.flowmap_composition_initial(f, 0.0, 2.0, 1.0, x, y, grid, p),
        lambda: nb.diagnostics.C_eig_2D(fm, 0.1, 0.1),
        lambda: nb.diagnostics.C_eig_aux_2D(fa, 0.1, 0.1),
        lambda: nb.diagnostics.C_tensor_2D(fa, 0.1, 0.1),
        lambda: nb.diagnostics.ftle_from_eig(np.ones((7, 6)), 2.0),
        lambda: nb.extraction.ftle_ridge_pts(np.ones((7, 6)), np.ones((7, 6, 2)), x, y),
        lambda: nb.extraction.ftle_ridges(np.ones((7, 6)), np.ones((7, 6, 2)), x, y),
        lambda: nb.extraction.percentile_value(np.ones(10), 50),
        lambda: nb.utils.binary_mask_dilation(np.zeros((7, 6), bool)),
        lambda: nb.flows.get_flow_linear_2D(((0, 1, 3), (0, 2, 7), (0, 1, 6)), np.zeros((3, 7, 6)), np.zeros((3, 7, 6))),
    ]
    for call in calls:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            call()


def test_argument_errors(lib):
    from numbacs_b200.flows import get_predefined_flow
    from numbacs_b200.integration import flowmap_grid_2D
    f, p, _ = get_predefined_flow("double_gyre")
    x, y = np.linspace(0, 2, 5), np.linspace(0, 1, 5)
    with pytest.raises(NotImplementedError):
        flowmap_grid_2D(f, 0.0, 8.0, x, y, p, method="lsoda")
    with pytest.raises(ValueError):
        flowmap_grid_2D(f, 0.0, 8.0, x, y, p, method="rk4")
    with pytest.raises(ValueError):
        flowmap_grid_2D(f, 0.0, 8.0, x, y, p, mask=np.zeros((3, 3), bool))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "numbacs_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn
                assert "oracle/" not in src.replace("never includes anything that lives under oracle/", ""), fn
