"""Multi-GPU test (needs >= 2 GPUs, skipped otherwise): the NCCL row-block path of
numbacs_b200.sharded gives the same bits as the single-GPU path."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nx, ny, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from numbacs_b200.flows import get_predefined_flow
    from numbacs_b200.sharded import flowmap_ftle_sharded, gather_rows
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    x = torch.linspace(0, 2, nx, dtype=torch.float64, device="cuda")
    y = torch.linspace(0, 1, ny, dtype=torch.float64, device="cuda")
    f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
    fm, ft, _ = flowmap_ftle_sharded(f, 0.0, -10.0, x, y, p, 2.0 / (nx - 1), 1.0 / (ny - 1))
    fm_all = gather_rows(fm.contiguous(), nx)
    ft_all = gather_rows(ft, nx)
    if rank == 0:
        np.savez(out_path, fm=fm_all.cpu().numpy(), ft=ft_all.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharding_matches_single_gpu(tmp_path, lib):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    from numbacs_b200.flows import get_predefined_flow
    from numbacs_b200.diagnostics import flowmap_ftle_grid_2D
    world = min(torch.cuda.device_count(), 4)
    nx, ny = 1026, 515
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(world, _free_port(), nx, ny, out), nprocs=world, join=True)
    got = np.load(out)
    f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
    x, y = np.linspace(0, 2, nx), np.linspace(0, 1, ny)
    # torch.linspace and np.linspace may differ in the last bit: use the same coordinates
    xt = torch.linspace(0, 2, nx, dtype=torch.float64).numpy()
    yt = torch.linspace(0, 1, ny, dtype=torch.float64).numpy()
    fm, ft = flowmap_ftle_grid_2D(f, 0.0, -10.0, xt, yt, p, 2.0 / (nx - 1), 1.0 / (ny - 1))
    assert np.array_equal(got["fm"], fm)
    assert np.array_equal(got["ft"], ft)


def _ridge_worker(rank, world, port, nx, ny, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from numbacs_b200.flows import get_predefined_flow
    from numbacs_b200.sharded import flowmap_ridges_sharded, gather_points, gather_rows
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    x = torch.linspace(0, 2, nx, dtype=torch.float64, device="cuda")
    y = torch.linspace(0, 1, ny, dtype=torch.float64, device="cuda")
    f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
    dx, dy = float(x[1] - x[0]), float(y[1] - y[0])
    fm, ft, pts, _ = flowmap_ridges_sharded(f, 0.0, -10.0, x, y, p, dx, dy, sdd_thresh=10.0)
    ft_all = gather_rows(ft.contiguous(), nx)
    pts_all = gather_points(pts)
    if rank == 0:
        np.savez(out_path, ft=ft_all.cpu().numpy(), pts=pts_all.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_multi_gpu_ridge_tail_matches_single_gpu(tmp_path, lib):
    """Config 5's tail on row blocks with a two-row NCCL halo == the single-GPU pipeline, bit for bit."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    from numbacs_b200.flows import get_predefined_flow
    from numbacs_b200.integration import flowmap_grid_2D
    from numbacs_b200.diagnostics import C_eig_2D, ftle_from_eig
    from numbacs_b200.extraction import ftle_ridge_pts
    world = min(torch.cuda.device_count(), 4)
    nx, ny = 1026, 515
    out = str(tmp_path / "ridges.npz")
    mp.spawn(_ridge_worker, args=(world, _free_port(), nx, ny, out), nprocs=world, join=True)
    got = np.load(out)
    f, p, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
    xt = torch.linspace(0, 2, nx, dtype=torch.float64).numpy()
    yt = torch.linspace(0, 1, ny, dtype=torch.float64).numpy()
    dx, dy = float(xt[1] - xt[0]), float(yt[1] - yt[0])
    fm = flowmap_grid_2D(f, 0.0, -10.0, xt, yt, p)
    vals, vecs = C_eig_2D(fm, dx, dy)
    ft = ftle_from_eig(vals[:, :, 1], -10.0)
    pts = ftle_ridge_pts(ft, vecs[:, :, :, 1], xt, yt, sdd_thresh=10.0)
    assert len(pts) > 100
    assert np.array_equal(got["ft"], ft)
    assert np.array_equal(got["pts"], pts)


def _lavd_fields():
    t, xs, ys = np.linspace(0, 10, 17), np.linspace(0, 2, 33), np.linspace(0, 1, 25)
    T, X, Y = np.meshgrid(t, xs, ys, indexing="ij")
    a = 0.25 * np.sin(0.2 * np.pi * T)
    b = 1 - 2 * a
    f = a * X ** 2 + b * X
    U = -np.pi * 0.1 * np.sin(np.pi * f) * np.cos(np.pi * Y)
    V = np.pi * 0.1 * np.cos(np.pi * f) * np.sin(np.pi * Y) * (2 * a * X + b)
    vort = np.sin(3 * X + 0.3 * T) * np.cos(2 * Y)
    return t, xs, ys, U, V, vort


def _lavd_worker(rank, world, port, nx, ny, n, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from numbacs_b200 import flows
    from numbacs_b200.sharded import gather_rows, lavd_flowmap_sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    t, xs, ys, U, V, vort = _lavd_fields()
    grid, Cu, Cv = flows.get_interp_arrays_2D(t, xs, ys, U, V)
    f = flows.get_flow_2D(grid, Cu, Cv, extrap_mode="linear")          # coefficients replicated per rank
    gw, Cw = flows.get_interp_arrays_scalar(t, xs, ys, vort)
    w = flows.get_callable_scalar(gw, Cw, extrap_mode="linear")
    x = torch.linspace(0.1, 1.9, nx, dtype=torch.float64, device="cuda")
    y = torch.linspace(0.1, 0.9, ny, dtype=torch.float64, device="cuda")
    lavd, ts, _ = lavd_flowmap_sharded(f, 1.0, 6.0, x, y, np.array([1.0]), w, n=n)
    full = gather_rows(lavd.contiguous(), nx)
    if rank == 0:
        np.savez(out_path, lavd=full.cpu().numpy(), ts=ts)
    dist.barrier()
    dist.destroy_process_group()


def test_multi_gpu_lavd_matches_single_gpu(tmp_path, lib):
    """Row-sharded fused LAVD with the NCCL all-reduce of the mean vorticity == single-GPU call."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    from numbacs_b200 import flows
    from numbacs_b200.diagnostics import lavd_flowmap_grid_2D
    world = min(torch.cuda.device_count(), 4)
    nx, ny, n = 203, 97, 41
    out = str(tmp_path / "lavd.npz")
    mp.spawn(_lavd_worker, args=(world, _free_port(), nx, ny, n, out), nprocs=world, join=True)
    got = np.load(out)
    t, xs, ys, U, V, vort = _lavd_fields()
    grid, Cu, Cv = flows.get_interp_arrays_2D(t, xs, ys, U, V)
    f = flows.get_flow_2D(grid, Cu, Cv, extrap_mode="linear")
    gw, Cw = flows.get_interp_arrays_scalar(t, xs, ys, vort)
    w = flows.get_callable_scalar(gw, Cw, extrap_mode="linear")
    x = torch.linspace(0.1, 1.9, nx, dtype=torch.float64).numpy()
    y = torch.linspace(0.1, 0.9, ny, dtype=torch.float64).numpy()
    ref, ts = lavd_flowmap_grid_2D(f, 1.0, 6.0, x, y, np.array([1.0]), w, n=n)
    assert np.array_equal(got["ts"], ts)
    # the mean is summed in a different order (per-rank partial sums): rounding-level agreement
    assert np.abs(got["lavd"] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
