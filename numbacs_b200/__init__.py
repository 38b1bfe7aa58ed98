"""numbacs_b200 -- B200-native flow-map + FTLE + LAVD hot path of NumbaCS.

Drop-in for the reference's hot-path API (same function names, arguments and array layouts):

    from numbacs_b200.flows import get_predefined_flow, get_interp_arrays_2D, get_flow_2D
    from numbacs_b200.integration import flowmap_grid_2D, flowmap_n_grid_2D, flowmap, flowmap_n
    from numbacs_b200.integration import flowmap_aux_grid_2D
    from numbacs_b200.diagnostics import ftle_grid_2D, lavd_grid_2D
    from numbacs_b200.diagnostics import C_tensor_2D, C_eig_aux_2D, C_eig_2D, ftle_from_eig
    from numbacs_b200.extraction import ftle_ridge_pts

Python only marshals pointers: all arithmetic runs in hand-written sm_100a CUDA kernels behind the
C-ABI of libb200cs.so (include/b200cs.h).  There is no CPU fallback.
"""
from . import _lib, diagnostics, extraction, flows, integration, utils  # noqa: F401

__version__ = "0.1.0"


def install_as_numbacs():
    """Alias this package as ``numbacs`` (numbacs.flows / .integration / .diagnostics) so that an
    unmodified reference script picks up the GPU path."""
    import sys
    sys.modules.setdefault("numbacs", sys.modules[__name__])
    for sub in ("flows", "integration", "diagnostics", "extraction", "utils"):
        sys.modules.setdefault("numbacs." + sub, sys.modules[__name__ + "." + sub])
