"""Row-block sharding of the flow-map + FTLE path across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for the host-logic tests).
The initial particle grid is split along axis 0 (x, the slowest axis of the C-order 'ij' layout,
so every block is contiguous); particles are independent, so the integration needs no
communication at all.  The FTLE stencil needs the flow map at rows i-1 / i+1
(/root/reference/src/numbacs/utils.py:41-44): each rank sends its first / last flow-map row
(ny x 2 float64 = 262 KB at ny = 16384) to its lower / upper neighbour -- one point-to-point
exchange, the only data-path collective -- and an optional gather assembles the result on rank 0.

Spline coefficient arrays are replicated: every rank creates its own flow handle.
"""
import numpy as np

__all__ = ["row_block", "exchange_halo_rows", "flowmap_ftle_sharded", "gather_rows"]


def row_block(nx, world_size, rank):
    """[i0, i1) of the rows owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(nx, world_size)
    i0 = rank * base + min(rank, rem)
    return i0, i0 + base + (1 if rank < rem else 0)


def exchange_halo_rows(slab, has_lo, has_hi, rank, group=None):
    """slab[(has_lo + rows + has_hi), ny, 2] with the owned rows already in place: sends the first
    / last OWNED row to rank-1 / rank+1 and receives their edge rows into slab[0] / slab[-1].
    Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch.distributed as dist
    ops = []
    if has_lo:
        ops.append(dist.P2POp(dist.isend, slab[1], rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, slab[0], rank - 1, group))
    if has_hi:
        ops.append(dist.P2POp(dist.isend, slab[-2], rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, slab[-1], rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def _cuda_backend():
    from .integration import flowmap_grid_2D
    from .diagnostics import ftle_slab_2D

    def integrate(funcptr, t0, T, x_rows, y, params, method, rtol, atol, out, info):
        # writes straight into the caller's slab rows (device tensor), no extra copy
        from . import _lib
        import ctypes as C
        xa, ya, pa = _lib.arg_in(x_rows), _lib.arg_in(y), _lib.arg_in(params)
        stats = info.setdefault("stats_dev", None)
        _lib.check(_lib.load().b200cs_flowmap_grid_2d(
            int(funcptr), float(t0), float(T), xa.ptr, int(xa.obj.shape[0]), ya.ptr,
            int(ya.obj.shape[0]), pa.ptr, int(pa.obj.shape[0]), _lib.METHOD_DOP853, float(rtol),
            float(atol), None, 0, C.c_void_p(out.data_ptr()), None, None, None,
            C.c_void_p(stats.data_ptr()) if stats is not None else None, _lib.current_stream(True)))

    def ftle(slab, T, dx, dy, halo):
        return ftle_slab_2D(slab, T, dx, dy, halo)

    return integrate, ftle


def flowmap_ftle_sharded(funcptr, t0, T, x, y, params, dx, dy, method="dop853", rtol=1e-6,
                         atol=1e-8, *, group=None, backend=None, info=None):
    """Each rank integrates its row block of the (x, y) grid and computes the FTLE rows it owns.

    Returns (flowmap_block [rows, ny, 2], ftle_block [rows, ny], (i0, i1)); tensors stay on the
    rank's device.  `backend` = (integrate, ftle) lets the CPU tests substitute the oracle for the
    CUDA library while exercising the same partition / halo / assembly logic."""
    import torch
    import torch.distributed as dist
    if method.lower() != "dop853":
        raise NotImplementedError("only method='dop853' is implemented on the GPU")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nx, ny = len(x), len(y)
    i0, i1 = row_block(nx, world, rank)
    rows = i1 - i0
    # ranks that own no rows (world > nx) take no part in the halo exchange
    has_lo = int(rank > 0 and rows > 0 and row_block(nx, world, rank - 1)[1] > row_block(nx, world, rank - 1)[0])
    has_hi = int(rank < world - 1 and rows > 0 and row_block(nx, world, rank + 1)[1] > row_block(nx, world, rank + 1)[0])
    integrate, ftle = backend if backend is not None else _cuda_backend()
    device = "cuda" if backend is None else "cpu"
    slab = torch.empty((has_lo + rows + has_hi, ny, 2), dtype=torch.float64, device=device)
    own = slab[has_lo:has_lo + rows]
    info = {} if info is None else info
    if rows:
        integrate(funcptr, t0, T, x[i0:i1], y, params, method, rtol, atol, own, info)
    if world > 1:
        exchange_halo_rows(slab, has_lo, has_hi, rank, group)
    ft = ftle(slab, T, dx, dy, (has_lo, has_hi)) if rows else torch.empty((0, ny), dtype=torch.float64, device=device)
    return own, ft, (i0, i1)


def gather_rows(block, nx, group=None, dst=0):
    """Gather row blocks [rows_r, ...] of every rank into the full [nx, ...] array on `dst`
    (None elsewhere).  Blocks may differ in size by one row, so they are padded to a common size."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return block
    max_rows = -(-nx // world)
    pad = torch.zeros((max_rows,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    pad[:block.shape[0]] = block
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        a, b = row_block(nx, world, r)
        parts.append(bufs[r][:b - a])
    return torch.cat(parts, dim=0)
