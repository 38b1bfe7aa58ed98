"""GPU tests of the lane mappings of the final-time flow-map kernels (numbacs_b200/csrc/flowmap_kernel.cuh).

Two launch shapes integrate the same particle with the same instruction sequence, so they must
agree BIT FOR BIT, step counts and statuses included:
  * the tiled one-particle-per-thread kernel (a warp = a kTileI x 32/kTileI tile of the grid), which
    small launches use, and the point-list kernel (a warp = 32 consecutive points);
  * the queue kernels (flowmap_init_kernel + flowmap_queue_kernel: pre-initialised particle slots,
    finished lanes fetch the next slot), which the Bickley jet uses from 65 536 particles upwards.
The reference for every case is the same particles pushed through `flowmap` in chunks of 30 000
points -- below the queue threshold, plain lane = point mapping (integration.py:7-61 is the
reference function both replace).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHUNK = 30000


@pytest.fixture(scope="module")
def nb(lib):
    import numbacs_b200 as nb
    from numbacs_b200 import _lib
    assert _lib.device_count() >= 1
    return nb


def by_chunks(nb, f, t0, T, pts, p):
    fm = np.empty_like(pts)
    steps = np.empty((len(pts), 2), np.int32)
    status = np.empty(len(pts), np.int32)
    nfev = 0
    for a in range(0, len(pts), CHUNK):
        info = {}
        fm[a:a + CHUNK] = nb.integration.flowmap(f, t0, T, pts[a:a + CHUNK], p, info=info)
        steps[a:a + CHUNK] = info["steps"]
        status[a:a + CHUNK] = info["status"]
        nfev += int(info["stats"][0])
    return fm, steps, status, nfev


@pytest.mark.parametrize("flow,nx,ny,T", [
    ("bickley_jet", 701, 203, 6.0),     # 142 303 particles: queue kernels, ragged 8 x 4 edge tiles
    ("bickley_jet", 333, 77, 6.0),      # 25 641: tiled one-particle-per-thread kernel
    ("double_gyre", 401, 201, -10.0),   # config 1: tiled kernel, ragged 4 x 8 edge tiles
])
def test_grid_launch_shapes_bit_identical(nb, flow, nx, ny, T):
    kw = {"int_direction": -1.0} if flow == "double_gyre" else {}
    f, p, dom = nb.flows.get_predefined_flow(flow, **kw)
    x = np.linspace(dom[0][0], dom[0][1], nx)
    y = np.linspace(-3.0, 3.0, ny) if flow == "bickley_jet" else np.linspace(dom[1][0], dom[1][1], ny)
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, T, x, y, p, info=info)
    X, Y = np.meshgrid(x, y, indexing="ij")
    pts = np.column_stack((X.ravel(), Y.ravel()))
    fm_c, steps_c, status_c, nfev_c = by_chunks(nb, f, 0.0, T, pts, p)
    assert np.array_equal(fm.reshape(-1, 2), fm_c)
    assert np.array_equal(np.asarray(info["steps"]).reshape(-1, 2), steps_c)
    assert np.array_equal(np.asarray(info["status"]).ravel(), status_c)
    assert int(info["stats"][0]) == nfev_c
    assert (status_c == 1).all()


def test_queue_kernel_masked_grid(nb):
    """Masked particles of a queue launch: zeros, MASKED status, no steps; the others as unmasked."""
    f, p, dom = nb.flows.get_predefined_flow("bickley_jet")
    x, y = np.linspace(dom[0][0], dom[0][1], 640), np.linspace(-3.0, 3.0, 150)   # 96 000 particles
    rng = np.random.default_rng(5)
    mask = rng.random((640, 150)) < 0.35
    mask[100:140] = True    # whole rows without work: slots the refill has to skip
    info = {}
    fm = nb.integration.flowmap_grid_2D(f, 0.0, 6.0, x, y, p, mask=mask, info=info)
    X, Y = np.meshgrid(x, y, indexing="ij")
    pts = np.column_stack((X[~mask], Y[~mask]))
    fm_c, steps_c, status_c, nfev_c = by_chunks(nb, f, 0.0, 6.0, pts, p)
    assert np.array_equal(fm[~mask], fm_c)
    assert (fm[mask] == 0.0).all()
    steps, status = np.asarray(info["steps"]), np.asarray(info["status"])
    assert np.array_equal(steps[~mask], steps_c) and (steps[mask] == 0).all()
    assert (status[mask] == 0).all() and (status[~mask] == 1).all()
    assert int(info["stats"][0]) == nfev_c


def test_queue_kernel_point_list(nb):
    """A point list above the queue threshold against the same points in small chunks."""
    f, p, dom = nb.flows.get_predefined_flow("bickley_jet")
    rng = np.random.default_rng(11)
    pts = np.column_stack((rng.uniform(dom[0][0], dom[0][1], 70001), rng.uniform(-3, 3, 70001)))
    info = {}
    fm = nb.integration.flowmap(f, 0.0, 6.0, pts, p, info=info)
    fm_c, steps_c, status_c, nfev_c = by_chunks(nb, f, 0.0, 6.0, pts, p)
    assert np.array_equal(fm, fm_c)
    assert np.array_equal(np.asarray(info["steps"]), steps_c)
    assert int(info["stats"][0]) == nfev_c


def test_queue_kernel_failed_particles(nb):
    """NaN particles in a queue launch end with a non-OK status and leave the others untouched."""
    f, p, dom = nb.flows.get_predefined_flow("bickley_jet")
    rng = np.random.default_rng(3)
    pts = np.column_stack((rng.uniform(dom[0][0], dom[0][1], 66000), rng.uniform(-3, 3, 66000)))
    bad = rng.choice(66000, 50, replace=False)
    ptsb = pts.copy()
    ptsb[bad, 0] = np.nan
    info, infob = {}, {}
    fm = nb.integration.flowmap(f, 0.0, 6.0, pts, p, info=info)
    fmb = nb.integration.flowmap(f, 0.0, 6.0, ptsb, p, info=infob)
    good = np.ones(66000, bool)
    good[bad] = False
    assert np.array_equal(fm[good], fmb[good])
    assert (np.asarray(infob["status"])[bad] != 1).all()
    assert (np.asarray(infob["status"])[good] == 1).all()
