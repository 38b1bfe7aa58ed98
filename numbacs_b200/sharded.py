"""Row-block sharding of the flow-map + FTLE path across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for the host-logic tests).
The initial particle grid is split along axis 0 (x, the slowest axis of the C-order 'ij' layout,
so every block is contiguous); particles are independent, so the integration needs no
communication at all.  The FTLE stencil needs the flow map at rows i-1 / i+1
(/root/reference/src/numbacs/utils.py:41-44): each rank sends its first / last flow-map row
(ny x 2 float64 = 262 KB at ny = 16384) to its lower / upper neighbour -- one point-to-point
exchange, the only data-path collective -- and an optional gather assembles the result on rank 0.
LAVD needs one all-reduce of n doubles (the spatial-mean vorticity per output time,
lavd_flowmap_sharded).  The ridge tail of config 5 (Cauchy-Green eigen-pairs -> FTLE -> ridge points) widens the halo to
two rows and is otherwise local (flowmap_ridges_sharded).

Spline coefficient arrays are replicated: every rank creates its own flow handle.
"""
import numpy as np

__all__ = ["row_block", "balanced_row_blocks", "estimate_row_cost", "exchange_halo_rows",
           "flowmap_ftle_sharded", "flowmap_ridges_sharded", "lavd_flowmap_sharded", "gather_rows",
           "gather_points"]


def row_block(nx, world_size, rank):
    """[i0, i1) of the rows owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(nx, world_size)
    i0 = rank * base + min(rank, rem)
    return i0, i0 + base + (1 if rank < rem else 0)


def balanced_row_blocks(cost_per_row, world_size, min_rows=1):
    """Contiguous row blocks [(i0, i1)] * world_size with (nearly) equal summed cost.

    Every rank gets at least `min_rows` rows (1 for the FTLE halo exchange, which talks to the
    immediate neighbour only; 2 for the ridge tail's two-row halo), so no empty block can sit
    between two non-empty ones; with fewer than world_size * min_rows rows the equal-size
    partition of row_block() is returned (empty blocks only at the end).

    The adaptive integrator does not spend the same time on every particle (7-23 step attempts on
    the double gyre), and the cost varies smoothly with x, so equal-sized row blocks leave the GPUs
    4-5 % out of balance at 8 ranks.  Boundaries are placed on the cumulative cost instead; every
    rank computes the same partition from the same (deterministic) cost estimate."""
    c = np.asarray(cost_per_row, dtype=np.float64)
    nx = len(c)
    if world_size <= 1 or nx == 0:
        return [(0, nx)] + [(nx, nx)] * (world_size - 1)
    cum = np.concatenate(([0.0], np.cumsum(np.maximum(c, 1e-300))))
    targets = cum[-1] * np.arange(1, world_size) / world_size
    cuts = np.searchsorted(cum, targets, side="left")
    # choose the nearer of the two candidate boundaries and keep the cuts monotone
    cuts = np.where((cuts > 0) & (np.abs(cum[np.maximum(cuts - 1, 0)] - targets) < np.abs(cum[np.minimum(cuts, nx)] - targets)),
                    cuts - 1, cuts)
    cuts = np.clip(np.maximum.accumulate(cuts), 0, nx)
    if nx < world_size * min_rows:
        return [row_block(nx, world_size, r) for r in range(world_size)]
    edges = [0] + [int(v) for v in cuts] + [nx]
    for r in range(1, world_size):          # forward pass: every block at least min_rows
        edges[r] = max(edges[r], edges[r - 1] + min_rows)
    for r in range(world_size - 1, 0, -1):  # backward pass: leave room for the blocks above
        edges[r] = min(edges[r], edges[r + 1] - min_rows)
    return [(edges[r], edges[r + 1]) for r in range(world_size)]


def estimate_row_cost(funcptr, t0, T, x, y, params, rtol=1e-6, atol=1e-8, rows=256, cols=256):
    """Step attempts per row of the (x, y) grid, estimated by integrating a rows x cols subsample
    (a fraction of a millisecond on the GPU) and interpolating along x.

    With CUDA tensors for x and y everything stays on the device -- the subsample is gathered, the
    per-particle step counts are reduced per row there, and only `rows` doubles come back -- so
    the whole planning pass costs about a millisecond and can sit inside an end-to-end step."""
    from .integration import flowmap_grid_2D
    nx, ny = len(x), len(y)
    ix = np.unique(np.linspace(0, nx - 1, min(rows, nx)).round().astype(np.int64))
    iy = np.unique(np.linspace(0, ny - 1, min(cols, ny)).round().astype(np.int64))
    info = {}
    if hasattr(x, "is_cuda") and x.is_cuda and hasattr(y, "is_cuda") and y.is_cuda:
        import torch
        xs = x[torch.as_tensor(ix, device=x.device)]
        ys = y[torch.as_tensor(iy, device=y.device)]
        flowmap_grid_2D(funcptr, t0, T, xs, ys, params, rtol=rtol, atol=atol, info=info, device_out=True)
        per_row = (info["steps"].sum(dim=(1, 2), dtype=torch.float64) / len(iy)).cpu().numpy()
    else:
        x = np.asarray(x.cpu() if hasattr(x, "cpu") else x, dtype=np.float64)
        y = np.asarray(y.cpu() if hasattr(y, "cpu") else y, dtype=np.float64)
        flowmap_grid_2D(funcptr, t0, T, x[ix], y[iy], params, rtol=rtol, atol=atol, info=info)
        per_row = np.asarray(info["steps"]).sum(axis=(1, 2)).astype(np.float64) / len(iy)
    return np.interp(np.arange(nx), ix, per_row)


def exchange_halo_rows(slab, has_lo, has_hi, rank, group=None, width=1):
    """slab[(w*has_lo + rows + w*has_hi), ny, 2] with the owned rows already in place: sends the
    first / last `width` OWNED rows to rank-1 / rank+1 and receives their edge rows into
    slab[:w] / slab[-w:].  width = 1 for the FTLE stencil, 2 for the ridge tail (the ridge stencil
    reads FTLE at i +- 1, which reads the flow map at i +- 2).  Works on CUDA tensors (NCCL) and
    CPU tensors (gloo)."""
    import torch.distributed as dist
    w = int(width)
    ops = []
    if has_lo:
        ops.append(dist.P2POp(dist.isend, slab[w:2 * w], rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, slab[:w], rank - 1, group))
    if has_hi:
        ops.append(dist.P2POp(dist.isend, slab[-2 * w:-w] if w > 0 else slab[:0], rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, slab[-w:], rank + 1, group))
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
                         atol=1e-8, *, group=None, backend=None, info=None, blocks=None):
    """Each rank integrates its row block of the (x, y) grid and computes the FTLE rows it owns.

    Returns (flowmap_block [rows, ny, 2], ftle_block [rows, ny], (i0, i1)); tensors stay on the
    rank's device.  `backend` = (integrate, ftle) lets the CPU tests substitute the oracle for the
    CUDA library while exercising the same partition / halo / assembly logic.  `blocks` is an
    optional list of (i0, i1) per rank (e.g. from balanced_row_blocks); default: equal sizes."""
    import torch
    import torch.distributed as dist
    if method.lower() != "dop853":
        raise NotImplementedError("only method='dop853' is implemented on the GPU")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nx, ny = len(x), len(y)
    if blocks is None:
        blocks = [row_block(nx, world, r) for r in range(world)]
    i0, i1 = blocks[rank]
    rows = i1 - i0
    # ranks that own no rows (world > nx) take no part in the halo exchange
    has_lo = int(rank > 0 and rows > 0 and blocks[rank - 1][1] > blocks[rank - 1][0])
    has_hi = int(rank < world - 1 and rows > 0 and blocks[rank + 1][1] > blocks[rank + 1][0])
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


def _cuda_ridge_backend():
    from .diagnostics import C_eig_2D, ftle_from_eig
    from .extraction import ftle_ridge_pts

    def ridge_tail(slab, T, dx, dy, x_slab, y, sdd_thresh):
        vals, vecs, ftle = C_eig_2D(slab, dx, dy, ftle_T=T)
        return ftle, ftle_ridge_pts(ftle, vecs[:, :, :, 1], x_slab, y, sdd_thresh=sdd_thresh, percentile=0,
                                    spacing=(dx, dy))

    return ridge_tail


def flowmap_ridges_sharded(funcptr, t0, T, x, y, params, dx, dy, sdd_thresh=0.0, method="dop853",
                           rtol=1e-6, atol=1e-8, *, group=None, backend=None, ridge_backend=None,
                           info=None, blocks=None):
    """Config 5 end to end on row blocks: each rank integrates its rows, exchanges TWO flow-map
    rows with each neighbour, and runs C_eig_2D -> ftle_from_eig -> ftle_ridge_pts on its slab.

    The ridge test at row i reads the FTLE field at i +- 1 and that reads the flow map at i +- 2,
    so with a two-row halo the whole tail is local; the ridge kernel's own border rule (rows
    2 .. m-3 of what it is given) then selects exactly the rows a rank owns (and, on the first /
    last rank, drops the two global border rows like the reference).  Ridge points of consecutive
    ranks concatenate to the single-process result (raveled pixel order).  percentile is fixed at
    0 (the example's setting): a global percentile would need a distributed selection.

    Returns (flowmap_block [rows, ny, 2], ftle_block [rows, ny], ridge_pts [k, 2], (i0, i1))."""
    import torch
    import torch.distributed as dist
    if method.lower() != "dop853":
        raise NotImplementedError("only method='dop853' is implemented on the GPU")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nx, ny = len(x), len(y)
    if blocks is None:
        blocks = [row_block(nx, world, r) for r in range(world)]
    i0, i1 = blocks[rank]
    rows = i1 - i0
    W = 2
    if world > 1 and min(b - a for a, b in blocks) < W:
        raise ValueError("every rank needs at least 2 rows for the two-row halo exchange")
    has_lo, has_hi = int(rank > 0), int(rank < world - 1)
    integrate, _ = backend if backend is not None else _cuda_backend()
    ridge_tail = ridge_backend if ridge_backend is not None else _cuda_ridge_backend()
    device = "cuda" if backend is None else "cpu"
    slab = torch.empty((W * has_lo + rows + W * has_hi, ny, 2), dtype=torch.float64, device=device)
    own = slab[W * has_lo:W * has_lo + rows]
    info = {} if info is None else info
    integrate(funcptr, t0, T, x[i0:i1], y, params, method, rtol, atol, own, info)
    if world > 1:
        exchange_halo_rows(slab, has_lo, has_hi, rank, group, width=W)
    x_slab = x[i0 - W * has_lo:i1 + W * has_hi]
    ftle_slab, pts = ridge_tail(slab, T, dx, dy, x_slab, y, sdd_thresh)
    # FTLE rows of the slab are valid from slab row 1 on (row 0 / m-1 lack a neighbour): the owned
    # rows are at offset 2*has_lo, except that global rows 0 and nx-1 are border rows (ftle = 0)
    return own, ftle_slab[W * has_lo:W * has_lo + rows], pts, (i0, i1)


def output_times(t0, T, n, p0):
    """tspan[k] = p0 * (p0 * linspace(t0, t0 + T, n))[k] exactly as flowmap_n_grid_2D returns it
    (integration.py:514, 533; numba linspace = start + k*step, last point forced to stop)."""
    step = ((t0 + T) - t0) / (n - 1)
    te = [p0 * (t0 + k * step) for k in range(n - 1)] + [p0 * (t0 + T)]
    return np.array([p0 * v for v in te], dtype=np.float64)


def _cuda_lavd_backend():
    from .diagnostics import lavd_flowmap_grid_2D, lavd_vort_sums

    def vort_sums(vort_interp, tspan, x_rows, y):
        import torch
        X, Y = torch.meshgrid(x_rows, y, indexing="ij")
        return lavd_vort_sums(vort_interp, tspan, X.reshape(-1), Y.reshape(-1), device_out=True)

    def lavd(funcptr, t0, T, x_rows, y, params, vort_interp, n, rtol, atol, px, py, vort_avg):
        out, _ = lavd_flowmap_grid_2D(funcptr, t0, T, x_rows, y, params, vort_interp, n=n, rtol=rtol,
                                      atol=atol, period_x=px, period_y=py, vort_avg=vort_avg,
                                      device_out=True)
        return out

    return vort_sums, lavd


def lavd_flowmap_sharded(funcptr, t0, T, x, y, params, vort_interp, n=50, method="dop853", rtol=1e-6,
                         atol=1e-8, period_x=0.0, period_y=0.0, *, group=None, backend=None,
                         blocks=None):
    """LAVD (flowmap_n_grid_2D + lavd_grid_2D, fused) on row blocks.  The only exchange is the
    spatial-mean vorticity per output time (diagnostics.py:324-331): every rank sums the vorticity
    over ITS rows of the initial grid, an all-reduce(sum) of n doubles makes the global means, and
    the fused trajectory kernel then runs without any further communication.

    Returns (lavd_block [rows, ny], tspan [n], (i0, i1)).  x, y: torch tensors (CUDA for the
    product path)."""
    import torch
    import torch.distributed as dist
    if method.lower() != "dop853":
        raise NotImplementedError("only method='dop853' is implemented on the GPU")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nx, ny = len(x), len(y)
    if blocks is None:
        blocks = [row_block(nx, world, r) for r in range(world)]
    i0, i1 = blocks[rank]
    vort_sums, lavd = backend if backend is not None else _cuda_lavd_backend()
    p0 = float(params[0])
    tspan = output_times(float(t0), float(T), int(n), p0)
    x_rows = x[i0:i1]
    if i1 > i0:
        sums = vort_sums(vort_interp, tspan, x_rows, y)
    else:
        sums = torch.zeros(n, dtype=torch.float64, device=x.device)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)   # n doubles: the only collective
    vort_avg = sums / float(nx * ny)
    if i1 > i0:
        out = lavd(funcptr, t0, T, x_rows, y, params, vort_interp, n, rtol, atol, period_x, period_y, vort_avg)
    else:
        out = torch.empty((0, ny), dtype=torch.float64, device=x.device)
    return out, tspan, (i0, i1)


def gather_points(pts, group=None, dst=0):
    """Concatenate per-rank point lists [k_r, 2] in rank order on `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return pts
    n = torch.tensor([pts.shape[0]], dtype=torch.int64, device=pts.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    kmax = int(max(int(c) for c in counts))
    pad = torch.zeros((max(kmax, 1), 2), dtype=pts.dtype, device=pts.device)
    pad[:pts.shape[0]] = pts
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:int(counts[r])] for r in range(world)], dim=0)


def gather_rows(block, nx, group=None, dst=0, blocks=None, out=None):
    """Assemble the row blocks [rows_r, ...] of every rank into the full [nx, ...] array on `dst`
    (None elsewhere): the final gather of the sharded path.

    `blocks` is the list of (i0, i1) per rank the blocks were computed with (balanced_row_blocks
    or the default row_block partition); sizes may differ arbitrarily.  The receiving rank posts
    one irecv per peer STRAIGHT INTO the rows of the assembled array (no padding, no staging copy,
    every peer's block arrives over its own NVLink path at the same time) and copies its own
    block; the other ranks post one isend.  `out` optionally supplies the [nx, ...] destination on
    `dst` (e.g. a buffer reused across frames)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if blocks is None:
        blocks = [row_block(nx, world, r) for r in range(world)]
    if len(blocks) != world or blocks[0][0] != 0 or blocks[-1][1] != nx or \
            any(blocks[r][1] != blocks[r + 1][0] for r in range(world - 1)):
        raise ValueError(f"blocks {blocks} do not tile [0, {nx}) over {world} ranks")
    i0, i1 = blocks[rank]
    if block.shape[0] != i1 - i0:
        raise ValueError(f"rank {rank} holds {block.shape[0]} rows, its block {blocks[rank]} has {i1 - i0}")
    if world == 1:
        if out is None:
            return block
        out.copy_(block)
        return out
    block = block.contiguous()
    ops = []
    if rank == dst:
        if out is None:
            out = torch.empty((nx,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
        for r in range(world):
            a, b = blocks[r]
            if r != dst and b > a:
                ops.append(dist.P2POp(dist.irecv, out[a:b], r, group))
    elif i1 > i0:
        ops.append(dist.P2POp(dist.isend, block, dst, group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    if rank == dst and i1 > i0:
        out[i0:i1].copy_(block)
    for req in reqs:
        req.wait()
    return out if rank == dst else None
