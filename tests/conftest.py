import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Frozen reference vectors (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure, never used by the product path)."""
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def lib():
    """libb200cs.so, built in-tree if stale."""
    from numbacs_b200 import _build, _lib
    _build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def strict(lib):
    """Context-manager factory that binds the Python API to libb200cs_strict.so, the parity-
    calibration build of the same sources (B200CS_STRICT, numbacs_b200/csrc/dop853.cuh): the
    reference's evaluation order with separately rounded operations and CUDA libm.  Its distance
    to the oracle is the floor the product build's distance is compared with."""
    from numbacs_b200 import _build, _lib
    _build.build_strict()
    return lambda: _lib.use_library(_lib.STRICT_LIB_PATH)


@pytest.fixture
def coords_dg():
    return np.linspace(0, 2, 21), np.linspace(0, 1, 11)


@pytest.fixture
def mask_dg():
    mask = np.zeros((21, 11), np.bool_)
    mask[::2, ::2] = True
    return mask
