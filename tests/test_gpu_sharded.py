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
