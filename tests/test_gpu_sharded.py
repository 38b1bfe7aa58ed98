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
