"""World-size 2 and 3 `gloo` tests of the multi-GPU host logic (row partition, one-row halo
exchange, slab assembly, gather) on CPU.  The compute backend is swapped for the CPU oracle (tests
may use it), so the result must be BIT-IDENTICAL to the single-process oracle: particles are
independent and the halo rows are exact copies."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_backend():
    import torch
    import oracle as O
    flow, _, _ = O.get_predefined_flow("double_gyre")

    def integrate(funcptr, t0, T, x_rows, y, params, method, rtol, atol, out, info):
        fm = O.flowmap_grid_2D(flow, t0, T, np.asarray(x_rows), np.asarray(y), np.asarray(params),
                               rtol=rtol, atol=atol)
        out.copy_(torch.from_numpy(fm))

    def ftle(slab, T, dx, dy, halo):
        nx = slab.shape[0]
        full = O.ftle_grid_2D(slab.numpy(), T, dx, dy)
        # emulate the slab semantics with the plain oracle: recompute halo-adjacent rows with the
        # neighbours present, drop the halo rows
        lo, hi = halo
        if lo or hi:
            # pad so that the oracle treats halo rows as interior neighbours
            pad = np.concatenate([slab.numpy()[:1]] * (1 if lo else 0) + [slab.numpy()] +
                                 [slab.numpy()[-1:]] * (1 if hi else 0), axis=0)
            full = O.ftle_grid_2D(pad, T, dx, dy)[(1 if lo else 0):pad.shape[0] - (1 if hi else 0)]
        return torch.from_numpy(np.ascontiguousarray(full[lo:nx - hi]))

    return integrate, ftle


def _worker(rank, world, port, nx, ny, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from numbacs_b200.sharded import flowmap_ftle_sharded, gather_rows, row_block
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    params = np.array([1.0, 0.1, 0.25, 0.0, 0.2 * np.pi, 0.0])
    fm, ft, (i0, i1) = flowmap_ftle_sharded(0, 0.0, 5.0, x, y, params, x[1] - x[0], y[1] - y[0],
                                            backend=_oracle_backend())
    assert (i0, i1) == row_block(nx, world, rank) and fm.shape[0] == i1 - i0 == ft.shape[0]
    fm_all = gather_rows(fm.contiguous(), nx)
    ft_all = gather_rows(ft, nx)
    if rank == 0:
        np.savez(out_path, fm=fm_all.numpy(), ft=ft_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _balanced_worker(rank, world, port, nx, ny, out_path):
    """flowmap_ftle_sharded + gather_rows on cost-balanced (unequal) row blocks: one block is far
    larger than ceil(nx / world), the case the equal-size gather used to break on."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from numbacs_b200.sharded import balanced_row_blocks, flowmap_ftle_sharded, gather_rows
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    params = np.array([1.0, 0.1, 0.25, 0.0, 0.2 * np.pi, 0.0])
    cost = np.where(np.arange(nx) < nx // 4, 20.0, 1.0)      # the first quarter is 20x as expensive
    blocks = balanced_row_blocks(cost, world)
    sizes = [b - a for a, b in blocks]
    assert max(sizes) > -(-nx // world) and min(sizes) >= 1, blocks
    fm, ft, (i0, i1) = flowmap_ftle_sharded(0, 0.0, 5.0, x, y, params, x[1] - x[0], y[1] - y[0],
                                            backend=_oracle_backend(), blocks=blocks)
    assert (i0, i1) == blocks[rank]
    fm_all = gather_rows(fm.contiguous(), nx, blocks=blocks)
    pre = torch.full((nx, ny), -1.0, dtype=torch.float64) if rank == 0 else None
    ft_all = gather_rows(ft, nx, blocks=blocks, out=pre)
    if rank == 0:
        assert ft_all is pre
        np.savez(out_path, fm=fm_all.numpy(), ft=ft_all.numpy())
    else:
        assert fm_all is None and ft_all is None
    # a block list that does not match what the rank holds is an error on every rank, not a hang
    with pytest.raises(ValueError):
        gather_rows(ft, nx, blocks=[(0, nx)] + [(nx, nx)] * (world - 1) if i1 - i0 != nx else [(0, 1)] * world)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_rows_balanced_blocks_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    import oracle as O
    nx, ny = 26, 9
    out = str(tmp_path / "bal.npz")
    mp.spawn(_balanced_worker, args=(world, _free_port(), nx, ny, out), nprocs=world, join=True)
    got = np.load(out)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    f, p, _ = O.get_predefined_flow("double_gyre")
    fm = O.flowmap_grid_2D(f, 0.0, 5.0, x, y, p)
    assert np.array_equal(got["fm"], fm)
    assert np.array_equal(got["ft"], O.ftle_grid_2D(fm, 5.0, x[1] - x[0], y[1] - y[0]))


def _oracle_ridge_backend():
    import torch
    import oracle as O

    def ridge_tail(slab, T, dx, dy, x_slab, y, sdd_thresh):
        vals, vecs = O.C_eig_2D(slab.numpy(), dx, dy)
        ftle = O.ftle_from_eig(vals[:, :, 1], T)
        pts = O.ftle_ridge_pts(ftle, vecs[:, :, :, 1], np.asarray(x_slab), np.asarray(y), sdd_thresh, 0,
                               spacing=(dx, dy))
        return torch.from_numpy(ftle), torch.from_numpy(np.ascontiguousarray(pts))

    return ridge_tail


def _ridge_worker(rank, world, port, nx, ny, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from numbacs_b200.sharded import flowmap_ridges_sharded, gather_points, gather_rows
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    params = np.array([1.0, 0.1, 0.25, 0.0, 0.2 * np.pi, 0.0])
    fm, ft, pts, (i0, i1) = flowmap_ridges_sharded(
        0, 0.0, 8.0, x, y, params, x[1] - x[0], y[1] - y[0], sdd_thresh=1.0,
        backend=_oracle_backend(), ridge_backend=_oracle_ridge_backend())
    ft_all = gather_rows(ft.contiguous(), nx)
    pts_all = gather_points(pts)
    if rank == 0:
        np.savez(out_path, ft=ft_all.numpy(), pts=pts_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _lavd_setup():
    import oracle as O
    t, xs, ys = np.linspace(0, 10, 9), np.linspace(0, 2, 17), np.linspace(0, 1, 13)
    T, X, Y = np.meshgrid(t, xs, ys, indexing="ij")
    vort = np.sin(3 * X + 0.3 * T) * np.cos(2 * Y)
    grid = ((t[0], t[-1], len(t)), (xs[0], xs[-1], len(xs)), (ys[0], ys[-1], len(ys)))
    w = O.get_callable_scalar_linear(grid, vort, extrap_mode="constant")
    flow, params, _ = O.get_predefined_flow("double_gyre")
    return O, flow, params, w


def _oracle_lavd_backend():
    import torch
    O, flow, _, w = _lavd_setup()

    def vort_sums(vort_interp, tspan, x_rows, y):
        X, Y = np.meshgrid(np.asarray(x_rows), np.asarray(y), indexing="ij")
        out = [w(np.column_stack((np.full(X.size, tk), X.ravel(), Y.ravel()))).sum() for tk in tspan]
        return torch.tensor(out, dtype=torch.float64)

    def lavd(funcptr, t0, T, x_rows, y, params, vort_interp, n, rtol, atol, px, py, vort_avg):
        fmn, ts = O.flowmap_n_grid_2D(flow, t0, T, np.asarray(x_rows), np.asarray(y), np.asarray(params), n=n,
                                      rtol=rtol, atol=atol)
        va = vort_avg.numpy()
        out = np.zeros(fmn.shape[:2])
        for i in range(fmn.shape[0]):
            for j in range(fmn.shape[1]):
                pts = np.column_stack((ts, fmn[i, j, :, 0], fmn[i, j, :, 1]))
                out[i, j] = O.composite_simpsons(np.abs(w(pts) - va), abs(ts[1] - ts[0]))
        return torch.from_numpy(out)

    return vort_sums, lavd


def _lavd_worker(rank, world, port, nx, ny, n, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from numbacs_b200.sharded import gather_rows, lavd_flowmap_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x, y = torch.linspace(0.1, 1.9, nx, dtype=torch.float64), torch.linspace(0.1, 0.9, ny, dtype=torch.float64)
    params = np.array([1.0, 0.1, 0.25, 0.0, 0.2 * np.pi, 0.0])
    lavd, ts, _ = lavd_flowmap_sharded(0, 1.0, 6.0, x, y, params, None, n=n, backend=_oracle_lavd_backend())
    full = gather_rows(lavd.contiguous(), nx)
    if rank == 0:
        np.savez(out_path, lavd=full.numpy(), ts=ts)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lavd_equals_single_process(tmp_path, oracle, world):
    """Row-sharded LAVD with the all-reduced spatial-mean vorticity == oracle.lavd_grid_2D on the
    whole grid (the mean is summed in a different order: agreement to rounding, not bits)."""
    import torch
    import torch.multiprocessing as mp
    nx, ny, n = 11, 7, 8
    out = str(tmp_path / "lavd.npz")
    mp.spawn(_lavd_worker, args=(world, _free_port(), nx, ny, n, out), nprocs=world, join=True)
    got = np.load(out)
    O, flow, params, w = _lavd_setup()
    x = torch.linspace(0.1, 1.9, nx, dtype=torch.float64).numpy()
    y = torch.linspace(0.1, 0.9, ny, dtype=torch.float64).numpy()
    fmn, ts = O.flowmap_n_grid_2D(flow, 1.0, 6.0, x, y, params, n=n)
    X, Y = np.meshgrid(x, y, indexing="ij")
    ref = O.lavd_grid_2D(fmn, ts, 6.0, w, X.ravel(), Y.ravel())
    assert np.array_equal(got["ts"], ts)
    assert np.abs(got["lavd"] - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("world,nx", [(2, 40), (3, 41)])
def test_sharded_ridge_tail_equals_single_process(tmp_path, oracle, world, nx):
    """Two-row halo: C_eig_2D -> ftle_from_eig -> ridge points per row block == single process."""
    import torch.multiprocessing as mp
    ny = 23
    out = str(tmp_path / "ridges.npz")
    mp.spawn(_ridge_worker, args=(world, _free_port(), nx, ny, out), nprocs=world, join=True)
    got = np.load(out)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    f, p, _ = oracle.get_predefined_flow("double_gyre")
    fm = oracle.flowmap_grid_2D(f, 0.0, 8.0, x, y, p)
    vals, vecs = oracle.C_eig_2D(fm, x[1] - x[0], y[1] - y[0])
    ft = oracle.ftle_from_eig(vals[:, :, 1], 8.0)
    pts = oracle.ftle_ridge_pts(ft, vecs[:, :, :, 1], x, y, 1.0, 0)
    assert len(pts) > 5
    assert np.array_equal(got["ft"], ft)
    assert np.array_equal(got["pts"], pts)


@pytest.mark.parametrize("world,nx", [(2, 24), (3, 23), (4, 3)])
def test_sharded_equals_single_process(tmp_path, oracle, world, nx):
    import torch.multiprocessing as mp
    ny = 13
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), nx, ny, out), nprocs=world, join=True)
    got = np.load(out)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    f, p, _ = oracle.get_predefined_flow("double_gyre")
    fm = oracle.flowmap_grid_2D(f, 0.0, 5.0, x, y, p)
    ft = oracle.ftle_grid_2D(fm, 5.0, x[1] - x[0], y[1] - y[0])
    assert np.array_equal(got["fm"], fm)
    assert np.array_equal(got["ft"], ft)


def test_row_block_partition():
    from numbacs_b200.sharded import row_block
    for nx in (0, 1, 7, 16, 16384, 16385):
        for world in (1, 2, 3, 8):
            blocks = [row_block(nx, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == nx
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_row_blocks():
    from numbacs_b200.sharded import balanced_row_blocks
    rng = np.random.default_rng(0)
    for nx in (8, 100, 16384):
        cost = 10 + 5 * np.sin(np.linspace(0, 6, nx)) + rng.random(nx)
        for world in (1, 2, 3, 8):
            blocks = balanced_row_blocks(cost, world)
            assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == nx
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            if nx >= 100:
                sums = np.array([cost[a:b].sum() for a, b in blocks])
                assert sums.max() / sums.mean() < 1.0 + 2.0 * cost.max() * world / cost.sum()
    # uniform cost -> (almost) equal sizes; degenerate inputs stay valid
    sizes = [b - a for a, b in balanced_row_blocks(np.ones(1000), 8)]
    assert max(sizes) - min(sizes) <= 1
    # fewer rows than ranks: the equal-size partition (empty blocks only at the end, never between)
    assert balanced_row_blocks(np.ones(2), 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert balanced_row_blocks(np.zeros(0), 3) == [(0, 0)] * 3
    # a cost concentrated in a few rows must not starve any rank of rows
    spike = np.ones(64)
    spike[:2] = 1e6
    for min_rows in (1, 2):
        blocks = balanced_row_blocks(spike, 8, min_rows=min_rows)
        assert blocks[0][0] == 0 and blocks[-1][1] == 64
        assert all(blocks[r][1] == blocks[r + 1][0] for r in range(7))
        assert min(b - a for a, b in blocks) >= min_rows, blocks
