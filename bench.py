#!/usr/bin/env python
"""bench.py -- headline benchmark: FTLE grid points/s for flow map + FTLE (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 16384]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE config 5 without the ridge tail -- double_gyre FTLE on a
16384 x 16384 grid, t0 = 0, T = -10, dop853, rtol 1e-6 / atol 1e-8, default flow parameters with
p[0] = -1 (README.md:101-120 of the reference at 41 x 82 times the particle count).  It fits one
GPU (6.4 GB of outputs), so N = 1 runs the whole grid and N > 1 row-shards the SAME grid
("scaling": "strong"): each rank integrates nx/N rows, exchanges one flow-map row with each
neighbour (NCCL point-to-point over NVLink) and computes its FTLE rows.

A step = one full flow map + FTLE pass.
  value : device-timed (CUDA events on the launching stream, barrier + synchronize on both sides,
          max over ranks), x / y already resident in HBM.
  e2e   : the same pass through the public API with HOST buffers: x, y copied host->device from
          pinned memory every step, the flow map AND the FTLE field copied device->host (24 B/point)
          into one host buffer per field that all ranks share, so rank 0 holds the assembled arrays;
          at N > 1 the planning pass (cost-balanced row cuts) runs inside every step.
  roofline : the integration kernel (FP64-FMA bound; neither HBM nor tensor), ALGORITHMIC flops
          F = N_att*346 + N_fev*35 + 40 per particle from the kernel's own step
          statistics, divided by the kernel time measured live with CUDA events; peak = FP64 DFMA
          peak measured in this run (MEASURED_PEAKS.json has no FP64 entry).
  roofline_ftle : the FTLE stencil kernel against the measured HBM copy bandwidth (24 B/pixel).
  ridge_tail    : config 5's "+ FTLE ridge extraction" (C_eig_2D -> ftle_from_eig -> ftle_ridge_pts on
          the step's flow map), timed on its own after the K steps and reported beside the metric.
  cpu_baseline  : the CPU oracle (a port of the reference algorithm, oracle/) on all host threads
          over a bounded contiguous row sample of the same grid (best of 3).
  parity        : the oracle's rows of that sample against the GPU's rows of the SAME 16384^2 grid
          (SURVEY section 8d "Parity gates"): step-count mismatches, max|dx|/L over step-matching
          particles and over all, FTLE relative L2 with the stencils touching a mismatch left out.
          At N > 1 the sample straddles the cut between rank 0's and rank 1's row blocks.
  gather_ms     : the final gather -- every rank's FTLE block into the assembled field on rank 0
          (sharded.gather_rows: one irecv per peer straight into the destination rows).

--impl reference times the reference's CPU implementation of the path: the UNMODIFIED reference
package (baseline/_ref/numbacs, copied there by oracle/install_reference.py) running over
oracle/shims, which supply the two third-party leaves that cannot be installed offline
(numbalsoda's dop853 as Hairer's DOP853 in C driven through the reference's own numba @cfunc,
interpolation.splines in numba) -- kind "reference-shim", numba threads = host cores.  The plain
C / OpenMP port of the same algorithm is reported beside it; it is the fallback (kind "port")
when the package or numba is missing.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

T0, TINT, RTOL, ATOL = 0.0, -10.0, 1e-6, 1e-8
F_STEP, F_RHS_DG, F_HINIT = 346, 35, 40


def workload_name(n):
    return f"double_gyre FTLE {n}x{n} grid, t0=0, T=-10, dop853, rtol=1e-6, atol=1e-8"


# ------------------------------------------------------------------ reference arm / CPU baseline

def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers; the
    CPU arm must not inherit that, so the oracle's thread count is set explicitly."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample(n, rows, reps=1, i0=None, keep=False):
    """Oracle flow map + FTLE on `rows` contiguous rows [i0, i0 + rows) of the n x n grid, all host
    threads, best of `reps`.  Returns a dict: points per second, seconds, threads and (keep=True)
    the flow map, per-particle step counts and FTLE rows for the parity block."""
    import oracle as O
    O.build()
    O.set_num_threads(host_threads())
    f, p, _ = O.get_predefined_flow("double_gyre", int_direction=-1.0)
    x, y = np.linspace(0, 2, n), np.linspace(0, 1, n)
    if i0 is None:
        i0 = n // 4
    xs = x[i0:i0 + rows]
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        fm, _, _, steps, _ = O.flowmap_grid_2D(f, T0, TINT, xs, y, p, rtol=RTOL, atol=ATOL, full=True)
        ft = O.ftle_grid_2D(fm, TINT, x[1] - x[0], y[1] - y[0])
        best = min(best, time.perf_counter() - t)
    out = {"pps": rows * n / best, "seconds": best, "threads": O.num_threads(), "rows": rows, "i0": i0}
    if keep:
        out.update(fm=fm, steps=steps, ftle=ft)
    return out


def pick_rows(n, target_s):
    """Rows of the n x n grid the oracle integrates in about target_s seconds (a multiple of 8)."""
    pps = cpu_sample(n, 8)["pps"]
    rows = int(pps * target_s / n) // 8 * 8
    return max(8, min(rows, n))


CPU_TARGET_S = 4.0   # seconds per repetition of the CPU sample, both arms (best of 3 -> ~12 s)


def reference_over_shims():
    """The UNMODIFIED reference package (baseline/_ref/numbacs, or /root/reference/src in the build
    container) over oracle/shims (numbalsoda -> Hairer DOP853 in C driven through the reference's
    own numba @cfunc; interpolation.splines in numba).  Returns the callables or None."""
    n_threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n_threads)      # torchrun exported 1: numba's omp layer would obey it
    os.environ["NUMBA_NUM_THREADS"] = str(n_threads)
    for src in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isdir(os.path.join(src, "numbacs")):
            break
    else:
        return None
    try:
        sys.path.insert(0, src)
        sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
        import numba
        from numbacs.flows import get_predefined_flow
        from numbacs.integration import flowmap_grid_2D
        from numbacs.diagnostics import ftle_grid_2D
        numba.set_num_threads(min(n_threads, numba.config.NUMBA_NUM_THREADS))
        return {"get_predefined_flow": get_predefined_flow, "flowmap_grid_2D": flowmap_grid_2D,
                "ftle_grid_2D": ftle_grid_2D, "threads": numba.get_num_threads(), "src": src}
    except Exception as exc:  # numba missing, import error: the C port remains
        sys.stderr.write(f"reference over shims unavailable: {exc!r}\n")
        return None


def ref_sample(R, n, rows, reps=1, i0=None):
    """The reference's own flowmap_grid_2D + ftle_grid_2D (README.md:107-120 call sequence) on `rows`
    contiguous rows of the n x n grid; best of `reps`."""
    f, p, _ = R["get_predefined_flow"]("double_gyre", int_direction=-1.0)
    x, y = np.linspace(0, 2, n), np.linspace(0, 1, n)
    if i0 is None:
        i0 = n // 4
    xs = np.ascontiguousarray(x[i0:i0 + rows])
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        fm = R["flowmap_grid_2D"](f, T0, TINT, xs, y, p)
        R["ftle_grid_2D"](fm, TINT, x[1] - x[0], y[1] - y[0])
        best = min(best, time.perf_counter() - t)
    return {"pps": rows * n / best, "seconds": best}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # one CPU run per node; the other ranks exit without work
    n = args.n
    R = reference_over_shims()
    if R is not None:
        ref_sample(R, n, 8)                               # JIT compilation, excluded
        pps = ref_sample(R, n, 8)["pps"]
        rows = args.cpu_rows or max(8, min(int(pps * CPU_TARGET_S / n) // 8 * 8, n))
        for _ in range(args.warmup):
            ref_sample(R, n, max(8, rows // 8))
        times = [ref_sample(R, n, rows)["seconds"] for _ in range(args.steps)]
        kind, threads = "reference-shim", R["threads"]
        how = ("the UNMODIFIED reference source (numbacs.integration.flowmap_grid_2D + numbacs.diagnostics."
               "ftle_grid_2D, its own numba prange loops and @cfunc right-hand side) over oracle/shims: "
               "numbalsoda.dop853 -> Hairer DOP853 in C called through the cfunc pointer, as numbalsoda does; "
               f"source {R['src']}")
    else:
        rows = args.cpu_rows or pick_rows(n, CPU_TARGET_S)
        for _ in range(args.warmup):
            cpu_sample(n, max(8, rows // 8))
        times = []
        for _ in range(args.steps):
            r = cpu_sample(n, rows)
            times.append(r["seconds"])
        kind, threads = "port", r["threads"]
        how = ("oracle/ C restatement of the reference algorithm (OpenMP over particles); the reference "
               "package is not available on this box (baseline/_ref missing or numba not importable)")
    t_step = float(np.mean(times))
    t_best = float(np.min(times))
    val = rows * n / t_step
    sample = f"{rows} contiguous rows x {n} columns of the {n}x{n} grid per step ({rows * n} particles)"
    # secondary figure: the plain C / OpenMP port of the same algorithm on the same rows (best of 3)
    port = cpu_sample(n, rows, reps=3) if kind != "port" else None
    line = {
        "impl": "reference", "metric": "FTLE grid points/s (flowmap+FTLE)", "value": val,
        "unit": "grid points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "grid points/s", "cores": threads, "kind": kind,
                         "sample": sample, "best_of_steps_value": rows * n / t_best, "note": how,
                         "c_port_same_rows": None if port is None else
                         {"value": port["pps"], "cores": port["threads"], "kind": "port",
                          "note": "oracle/ C + OpenMP restatement, no numba, no callback indirection: the best "
                                  "this CPU does with the identical algorithm; NOT the reference"}},
        "e2e": {"value": val, "unit": "grid points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.tmp.read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.tmp.name)
        if sm:
            # the load samples are the upper half (idle samples before/after the region are lower power)
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


# ------------------------------------------------------------------ the B200 arm

def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the host pages
    it first-touches -- its slice of the shared result buffers -- are allocated next to the PCIe root
    the GPU writes through.  Returns the node or None."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out            # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def shared_pinned(name, count, create, own=None):
    """A float64 host buffer of `count` elements that every rank of the node maps (POSIX shared
    memory), so each GPU downloads its rows of the assembled result straight into its slice -- the
    final gather to host memory uses all PCIe links at once and rank 0 ends up holding the whole
    field.  `own` = (first, last) element range this rank writes: it is first-touched here (pages
    land on this process's NUMA node) and only that range is registered with CUDA.
    Returns (tensor, path, registered_ptr) or (None, path, None)."""
    import torch
    path = f"/dev/shm/{name}"
    try:
        if create:
            with open(path, "wb") as fh:
                fh.truncate(count * 8)
        t = torch.from_file(path, shared=True, size=count, dtype=torch.float64)
        lo, hi = (0, count) if own is None else (max(0, own[0]), min(count, own[1]))
        page = 4096 // 8
        lo, hi = lo // page * page, min(count, -(-hi // page) * page)
        if hi > lo:
            t[lo:hi].zero_()                                   # first touch
        ptr = t.data_ptr() + lo * 8
        rc = torch.cuda.cudart().cudaHostRegister(ptr, (hi - lo) * 8, 0) if hi > lo else 0
        if int(rc) != 0:
            return None, path, None
        return t, path, ptr
    except Exception:
        return None, path, None


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from numbacs_b200 import _build, _lib
    from numbacs_b200.flows import get_predefined_flow
    from numbacs_b200.sharded import (balanced_row_blocks, estimate_row_cost, exchange_halo_rows,
                                      gather_rows, gather_points)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not os.path.exists(_lib.LIB_PATH):
        _build.build()
    L = _lib.load()
    if _lib.device_count() < 1:
        raise RuntimeError("no CUDA device: bench.py --impl b200 has no CPU fallback")
    torch.cuda.set_device(local)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner to stdout (fd 1) when NCCL_DEBUG=VERSION/INFO is set in the
        # environment; the contract is ONE JSON line on stdout, so fd 1 points at stderr until the
        # line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    K, W = args.steps, args.warmup
    dx, dy = 2.0 / (n - 1), 1.0 / (n - 1)
    HW = 2 if world > 1 else 0     # halo rows per interior edge: 2 (the ridge tail reads i +- 2; FTLE needs 1)
    f, params, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
    x_host = torch.linspace(0, 2, n, dtype=torch.float64).pin_memory()
    y_host = torch.linspace(0, 1, n, dtype=torch.float64).pin_memory()
    x_dev, y_dev = x_host.cuda(), y_host.cuda()

    # row blocks of equal estimated COST (step attempts), not equal size: a 256 x 256 subsample of
    # the grid is integrated on the device and reduced per row there (every rank gets the same cuts)
    def plan():
        if world == 1:
            return [(0, n)]
        cost = estimate_row_cost(f, T0, TINT, x_dev, y_dev, params, RTOL, ATOL)
        return balanced_row_blocks(cost, world, min_rows=HW)

    plan()
    torch.cuda.synchronize()
    t_plan = []
    for _ in range(5):
        t = time.perf_counter()
        blocks = plan()
        torch.cuda.synchronize()
        t_plan.append(time.perf_counter() - t)
    plan_ms = float(np.median(t_plan)) * 1e3
    i0, i1 = blocks[rank]
    rows = i1 - i0
    has_lo, has_hi = int(rank > 0), int(rank < world - 1)

    x_rows_dev = x_dev[i0:i1].contiguous()
    slab = torch.empty((HW * has_lo + rows + HW * has_hi, n, 2), dtype=torch.float64, device="cuda")
    own = slab[HW * has_lo:HW * has_lo + rows]
    # the FTLE kernel sees the slab with ONE halo row per interior edge
    ftle_view = slab[(HW - 1) * has_lo:slab.shape[0] - (HW - 1) * has_hi]
    ftle = torch.empty((rows, n), dtype=torch.float64, device="cuda")
    stats = torch.zeros(3, dtype=torch.int64, device="cuda")
    p_arr = np.ascontiguousarray(params)
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)

    def k_flowmap(xr, yv, with_stats):
        _lib.check(L.b200cs_flowmap_grid_2d(
            f, T0, TINT, C.c_void_p(xr.data_ptr()), rows, C.c_void_p(yv.data_ptr()), n,
            C.c_void_p(p_arr.ctypes.data), len(p_arr), 0, RTOL, ATOL, None, 0,
            C.c_void_p(own.data_ptr()), None, None, None,
            C.c_void_p(stats.data_ptr()) if with_stats else None, sptr))

    def k_ftle():
        _lib.check(L.b200cs_ftle_slab_2d(C.c_void_p(ftle_view.data_ptr()), ftle_view.shape[0], n, TINT, dx, dy,
                                         None, has_lo, has_hi, C.c_void_p(ftle.data_ptr()), sptr))

    def halo():
        if world > 1:
            exchange_halo_rows(slab, has_lo, has_hi, rank, width=HW)

    def step(xr, yv, with_stats=False):
        k_flowmap(xr, yv, with_stats)
        halo()
        k_ftle()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = _lib.fp64_peak(20000)[0]          # measured DFMA peak of this GPU, TFLOP/s

    # ---- warm-up (also collects the step statistics once)
    stats.zero_()
    step(x_rows_dev, y_dev, with_stats=True)
    for _ in range(max(W - 1, 0)):
        step(x_rows_dev, y_dev)
    sync_all()
    st = stats.cpu().numpy().astype(np.float64)   # sum nfev, accepted, rejected of this rank's block

    # ---- device-timed region
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sync_all()
    t_wall = time.perf_counter()
    for k in range(K):
        ev[k][0].record(stream)
        k_flowmap(x_rows_dev, y_dev, False)
        ev[k][1].record(stream)
        halo()
        ev[k][2].record(stream)
        k_ftle()
        ev[k][3].record(stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall
    # back-to-back steps on one stream: the span first->last event covers exactly K steps
    span_ms = ev[0][0].elapsed_time(ev[K - 1][3])
    fm_ms = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(K)]))
    halo_ms = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(K)]))  # incl. waiting for the neighbour
    ft_ms = float(np.mean([ev[k][2].elapsed_time(ev[k][3]) for k in range(K)]))
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None

    # ---- the final gather over NVLink: every rank's FTLE block into the assembled field on rank 0
    gather_ms = 0.0
    if world > 1:
        full_ftle = torch.empty((n, n), dtype=torch.float64, device="cuda") if rank == 0 else None
        gather_rows(ftle, n, blocks=blocks, out=full_ftle)       # warm-up (NCCL connections)
        sync_all()
        best = float("inf")
        for _ in range(3):
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            gather_rows(ftle, n, blocks=blocks, out=full_ftle)
            g1.record(stream)
            torch.cuda.synchronize()
            best = min(best, g0.elapsed_time(g1))
            dist.barrier()
        gather_ms = best
        gathered_checksum = float(full_ftle.sum()) if rank == 0 else 0.0
        del full_ftle
        torch.cuda.empty_cache()

    # ---- end-to-end region: per step, every rank re-plans its row block (the 256 x 256 cost
    # estimate, on the device) and makes ONE C-ABI call with HOST buffers only -- x / y in from
    # pinned memory, its FLOW-MAP rows and its FTLE rows out to host memory
    # (b200cs_flowmap_ftle_grid_2d integrates in row chunks and streams finished rows over PCIe
    # while the next chunk is integrated).  The host destination is ONE buffer per field shared by
    # all ranks of the node (POSIX shared memory, registered with CUDA in every process): each GPU
    # writes its rows into its slice over its own PCIe link, so when the step ends rank 0 holds the
    # assembled [n, n, 2] flow map and [n, n] FTLE field -- what the reference's flowmap_grid_2D +
    # ftle_grid_2D return -- without any device-side gather.  Multi-GPU: every rank passes its row
    # block plus one stencil-only halo row per interior edge (recomputed, 2 of nx/N rows), so the
    # end-to-end path needs no exchange at all.
    tag = f"b200cs_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}"
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    # element ranges this rank writes: its FTLE rows; its flow-map rows plus one halo row per interior edge
    own_ft = (i0 * n, i1 * n)
    own_fm = ((i0 - has_lo) * 2 * n, (i1 + has_hi) * 2 * n)
    if world > 1:
        if rank == 0:
            ft_sh, ft_path, ft_reg = shared_pinned(tag + "_ftle", n * n, True, own_ft)
            fm_sh, fm_path, fm_reg = shared_pinned(tag + "_fm", 2 * n * n, True, own_fm)
        dist.barrier()
        if rank != 0:
            ft_sh, ft_path, ft_reg = shared_pinned(tag + "_ftle", n * n, False, own_ft)
            fm_sh, fm_path, fm_reg = shared_pinned(tag + "_fm", 2 * n * n, False, own_fm)
        ok = torch.tensor([int(ft_sh is not None and fm_sh is not None)], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        assembled = bool(int(ok))
    else:
        ft_sh, ft_path, ft_reg = shared_pinned(tag + "_ftle", n * n, True)
        fm_sh, fm_path, fm_reg = shared_pinned(tag + "_fm", 2 * n * n, True)
        assembled = ft_sh is not None and fm_sh is not None
    if not assembled:   # no shared memory: every rank keeps its rows in its own pinned buffer
        ft_reg = fm_reg = None
        ft_sh = torch.empty(n * n if world == 1 else rows * n, dtype=torch.float64).pin_memory()
        fm_sh = torch.empty(2 * (n * n if world == 1 else rows * n), dtype=torch.float64).pin_memory()
    ft_full = ft_sh.view(-1, n)
    fm_full = fm_sh.view(-1, n, 2)

    def e2e_step():
        b = plan()
        a0, a1 = b[rank]
        r0 = a0 if assembled or world == 1 else 0
        xs = x_host[a0 - has_lo:a1 + has_hi]
        _lib.check(L.b200cs_flowmap_ftle_grid_2d(
            f, T0, TINT, C.c_void_p(xs.data_ptr()), xs.shape[0], C.c_void_p(y_host.data_ptr()), n,
            C.c_void_p(p_arr.ctypes.data), len(p_arr), 0, RTOL, ATOL, None, dx, dy, has_lo, has_hi,
            None, C.c_void_p(ft_full[r0:].data_ptr()), None, None, sptr))
        return a0, a1

    def e2e_step_fm():
        # the same call, the flow-map rows downloaded as well (halo rows land in a scratch slab:
        # the C entry writes flowmap_out for every row it integrates)
        b = plan()
        a0, a1 = b[rank]
        r0 = a0 if assembled or world == 1 else 0
        xs = x_host[a0 - has_lo:a1 + has_hi]
        dst = fm_full[r0 - has_lo:] if (assembled and world > 1) else fm_full
        _lib.check(L.b200cs_flowmap_ftle_grid_2d(
            f, T0, TINT, C.c_void_p(xs.data_ptr()), xs.shape[0], C.c_void_p(y_host.data_ptr()), n,
            C.c_void_p(p_arr.ctypes.data), len(p_arr), 0, RTOL, ATOL, None, dx, dy, has_lo, has_hi,
            C.c_void_p(dst.data_ptr()), C.c_void_p(ft_full[r0:].data_ptr()), None, None, sptr))

    def timed_e2e(fn):
        fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
        return ms

    e2e_ms = timed_e2e(e2e_step)
    # halo rows of neighbouring ranks overlap by one row in the shared flow-map buffer (both write
    # the same values), so the with-flow-map variant is only run when that is well defined
    e2e_fm_ms = timed_e2e(e2e_step_fm) if (world == 1 or assembled) else None
    sync_all()
    if assembled or world == 1:
        checksum = float(ft_full.sum()) if rank == 0 else 0.0   # the assembled host field, read on rank 0
    else:
        checksum = float(ft_full[:rows].sum())

    # ---- parity block: oracle rows vs the GPU's rows of the same grid
    parity = None
    cpu = None
    if not args.no_cpu:
        parity, cpu = parity_block(args, n, world, rank, blocks, slab, own, ftle, HW, has_lo, f, params,
                                   x_dev, y_dev, dx, dy)

    # ---- config 5's tail, reported beside the metric (not inside it): Cauchy-Green eigen-pairs ->
    # FTLE from the largest eigenvalue -> sub-pixel ridge points, on the flow map of the last step
    # (examples/ftle/plot_dg_ftle_ridges.py:43-72: percentile 0, sdd_thresh 10).  At N > 1 every
    # rank runs the tail on its slab (two-row halo, sharded.flowmap_ridges_sharded's scheme) and
    # the ridge points are gathered on rank 0.
    ridge = None
    if not args.no_ridges:
        from numbacs_b200.diagnostics import C_eig_2D, ftle_from_eig
        from numbacs_b200.extraction import ftle_ridge_pts
        x_slab = x_dev[i0 - HW * has_lo:i1 + HW * has_hi]

        def tail():
            vals, vecs, ft2 = C_eig_2D(slab, dx, dy, ftle_T=TINT)
            return ft2, ftle_ridge_pts(ft2, vecs[:, :, :, 1], x_slab, y_dev, sdd_thresh=10.0, percentile=0,
                                       spacing=(dx, dy))

        tail()
        sync_all()
        best = float("inf")
        for _ in range(3):
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            ft2, rp = tail()
            r1.record(stream)
            torch.cuda.synchronize()
            best = min(best, r0.elapsed_time(r1))
        ft2_own = ft2[HW * has_lo:HW * has_lo + rows]
        num = float(torch.linalg.norm(ft2_own - ftle) ** 2)
        den = float(torch.linalg.norm(ftle) ** 2)
        tt = torch.tensor([best, float(rp.shape[0]), num, den], dtype=torch.float64, device="cuda")
        g_ms = 0.0
        if world > 1:
            mx = tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            best = float(mx[0])
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            allp = gather_points(rp, dst=0)
            g1.record(stream)
            torch.cuda.synchronize()
            g_ms = g0.elapsed_time(g1)
            del allp
        ridge = {"ms": best, "n_ridge_pts": int(tt[1]),
                 "ftle_vs_ftle_grid_2D_rel_l2": float(np.sqrt(float(tt[2]) / float(tt[3]))),
                 "gather_points_ms": g_ms,
                 "kernels": "C_eig_2D with the FTLE fused in + ftle_ridge_pts (one-evaluation detect with a bit mask, scan, compact), device-timed, "
                            "best of 3, max over ranks; sdd_thresh=10, percentile=0",
                 "algorithmic_bytes_per_point": 64 + 16 + 2 * 24}
        del ft2, rp
        torch.cuda.empty_cache()

    # ---- max over ranks
    t = torch.tensor([span_ms, e2e_ms, fm_ms, ft_ms, halo_ms, gather_ms, plan_ms,
                      e2e_fm_ms if e2e_fm_ms is not None else 0.0], dtype=torch.float64, device="cuda")
    agg = torch.cat([torch.tensor(st, dtype=torch.float64, device="cuda"),
                     torch.tensor([checksum], dtype=torch.float64, device="cuda")])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    span_ms, e2e_ms, fm_ms, ft_ms, halo_ms, gather_ms, plan_ms, e2e_fm_max = t.tolist()
    nfev, nacc, nrej, checksum = agg.tolist()
    if e2e_fm_ms is not None:
        e2e_fm_ms = e2e_fm_max

    # release the shared host buffers
    try:
        if assembled:
            torch.cuda.cudart().cudaHostUnregister(ft_reg)
            torch.cuda.cudart().cudaHostUnregister(fm_reg)
        del ft_full, fm_full, ft_sh, fm_sh
        if world > 1:
            dist.barrier()
        if rank == 0:
            for pth in (ft_path, fm_path):
                if pth and os.path.exists(pth):
                    os.unlink(pth)
    except Exception:
        pass

    if rank == 0:
        pts = float(n) * n
        ms_per_step = span_ms / K
        value = pts / (ms_per_step * 1e-3)
        e2e_val = pts / (e2e_ms / K * 1e-3)
        # algorithmic flops of the integration (whole job) and per-rank kernel rate
        n_att = nacc + nrej
        flops = n_att * F_STEP + nfev * F_RHS_DG + pts * F_HINIT
        fm_tflops_per_gpu = flops / world / (fm_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(f"flowmap_dg_{n}_world{world}")
            ftle_traffic = tr.get(f"ftle_{n}_world{world}")
        except Exception:
            ftle_traffic = None
        ftle_gbs = 24.0 * pts / world / (ft_ms * 1e-3) / 1e9
        # end to end = what flowmap_grid_2D + ftle_grid_2D hand back to the caller: the [n, n, 2] flow map
        # AND the [n, n] FTLE field in host memory (24 B/point over PCIe); the FTLE-only variant
        # (flow map left in HBM) is reported beside it
        h2d = int(8 * (n + 2 * (world - 1) + n * world))
        how = ("one b200cs_flowmap_ftle_grid_2d call per rank and step, host pointers only: x/y from pinned "
               "memory, flow-map and FTLE rows into ONE host buffer per field shared by all ranks (POSIX shared "
               "memory registered with CUDA in every process), so rank 0 holds the assembled arrays when the "
               "step ends; downloads overlap the integration of later row chunks; the per-step planning pass "
               "(cost-balanced cuts) is inside the timed region")
        ftle_only = {"value": e2e_val, "ms_per_step": e2e_ms / K, "d2h_bytes_per_step": int(8 * pts),
                     "note": "flowmap_out = NULL: only the FTLE field leaves the GPU"}
        if e2e_fm_ms is not None:
            e2e_entry = {"value": pts / (e2e_fm_ms / K * 1e-3), "unit": "grid points/s", "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": int(24 * pts), "ms_per_step": e2e_fm_ms / K,
                         "assembled_on_rank0_host": True, "planning_inside": world > 1, "result": how,
                         "numa_node_of_rank0": numa_node, "ftle_only": ftle_only}
        else:
            e2e_entry = {"value": e2e_val, "unit": "grid points/s", "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": int(8 * pts), "ms_per_step": e2e_ms / K,
                         "assembled_on_rank0_host": False, "planning_inside": world > 1,
                         "result": how + " [no shared memory on this box: every rank kept its FTLE rows in its own "
                                         "pinned buffer and the flow map stayed in HBM]"}
        line = {
            "metric": "FTLE grid points/s (flowmap+FTLE)", "value": value, "unit": "grid points/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "parallelism": f"row-block x{world}" + (", blocks balanced by estimated step count "
                                                              f"(planning {plan_ms:.2f} ms per call: outside the "
                                                              "device-timed steps, INSIDE every end-to-end step)"
                                                              if world > 1 else ""),
                       "row_blocks": [list(b) for b in blocks],
                       "l2": "no flush: every step rewrites 24 B/point of outputs "
                             f"({24 * pts / world / 1e9:.2f} GB per GPU >> 126 MB L2), inputs are 2 x {n} doubles"},
            "e2e": e2e_entry,
            "gpu_launches": 2 * K * world,  # timed (device) region: flow-map + FTLE kernel per step and rank
            "roofline": {"bound": "fp64", "achieved": fm_tflops_per_gpu, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": fm_tflops_per_gpu / fp64_peak, "traffic": traffic,
                         "kernel": "flowmap_kernel<DoubleGyre>", "kernel_ms": fm_ms,
                         "peak_source": "FP64 DFMA chain measured in this run (b200cs_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_particle": flops / pts, "nfev_per_particle": nfev / pts,
                         "attempts_per_particle": n_att / pts,
                         "note": "integration is FP64-FMA bound (no dense contraction, 16 B/particle of HBM "
                                 "traffic); ncu pipe utilisation is in profiles/"},
            "roofline_ftle": {"bound": "hbm", "achieved": ftle_gbs, "peak": hbm_peak, "unit": "GB/s",
                              "frac": ftle_gbs / hbm_peak, "traffic": ftle_traffic, "kernel": "ftle_kernel",
                              "kernel_ms": ft_ms, "peak_source": hbm_src, "bytes_per_point": 24},
            "parity": parity,
            "ridge_tail": ridge,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "halo_exchange_ms": halo_ms if world > 1 else 0.0,
            "gather_ms": gather_ms,
            "planning_ms": plan_ms if world > 1 else 0.0,
            "wall_ms_per_step": t_wall / K * 1e3,
            "ftle_checksum": checksum,
            "fp64_peak_tflops_measured": fp64_peak,
            "library": _lib.library_info(),
        }
        if world > 1:
            line["gather"] = {"ms": gather_ms, "bytes_into_rank0": int(8 * (n - (blocks[0][1] - blocks[0][0])) * n),
                              "GBps": 8.0 * (n - (blocks[0][1] - blocks[0][0])) * n / (gather_ms * 1e-3) / 1e9,
                              "checksum": gathered_checksum,
                              "how": "sharded.gather_rows: one NCCL irecv per peer straight into the rows of the "
                                     "assembled [n, n] FTLE field on rank 0, best of 3, max over ranks"}
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_block(args, n, world, rank, blocks, slab, own, ftle, HW, has_lo, f, params, x_dev, y_dev, dx, dy):
    """Oracle rows against the GPU's rows of the same grid (rank 0 reports).  The sample is a
    contiguous row range: at N = 1 the cpu_baseline sample; at N > 1 a shorter one that straddles
    the cut between rank 0 and rank 1, whose upper half rank 1 sends over."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity_common import compare_flowmaps, ftle_rel_l2
    from numbacs_b200.integration import flowmap_grid_2D
    from numbacs_b200.diagnostics import ftle_grid_2D
    if world == 1:
        rows_s = args.cpu_rows or pick_rows(n, CPU_TARGET_S)
        a = n // 4
    else:
        rows_s = min(args.cpu_rows or 256, 2 * min(blocks[0][1] - blocks[0][0], blocks[1][1] - blocks[1][0]))
        rows_s = max(8, rows_s // 2 * 2)
        a = blocks[0][1] - rows_s // 2
    b = a + rows_s
    # the GPU's rows [a, b) of the step's flow map and FTLE field, assembled on rank 0
    gpu_fm = gpu_ft = None
    if world == 1:
        gpu_fm, gpu_ft = own[a:b], ftle[a:b]
    else:
        cut = blocks[0][1]
        if rank == 0:
            hi_fm = torch.empty((b - cut, n, 2), dtype=torch.float64, device="cuda")
            hi_ft = torch.empty((b - cut, n), dtype=torch.float64, device="cuda")
            for req in dist.batch_isend_irecv([dist.P2POp(dist.irecv, hi_fm, 1), dist.P2POp(dist.irecv, hi_ft, 1)]):
                req.wait()
            gpu_fm = torch.cat([own[a:cut], hi_fm])
            gpu_ft = torch.cat([ftle[a:cut], hi_ft])
        elif rank == 1:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, own[:b - cut].contiguous(), 0),
                                               dist.P2POp(dist.isend, ftle[:b - cut].contiguous(), 0)]):
                req.wait()
    parity = cpu = None
    if rank == 0:
        reps = 3 if world == 1 else 1
        r = cpu_sample(n, rows_s, reps=reps, i0=a, keep=True)
        # per-particle step counts of the GPU for the sample rows: one extra (untimed) launch on
        # those rows; its positions must be bit-identical to the rows of the timed full-grid launch
        info = {}
        fm_s = flowmap_grid_2D(f, T0, TINT, x_dev[a:b].contiguous(), y_dev, params, rtol=RTOL, atol=ATOL,
                               info=info, device_out=True)
        identical = bool(torch.equal(fm_s, gpu_fm))
        st = info["steps"].cpu().numpy()
        res, same = compare_flowmaps(gpu_fm.cpu().numpy(), st, r["fm"], r["steps"], (2.0, 1.0))
        # FTLE rows a+1 .. b-2 have their full stencil inside the sample
        ft_gpu = gpu_ft[1:-1].cpu().numpy()
        keep_rows = slice(1, rows_s - 1)
        # the oracle's FTLE of the block treats its first / last row as borders; compare interior rows
        bad = ~same
        touch = bad.copy()
        touch[1:] |= bad[:-1]
        touch[:-1] |= bad[1:]
        touch[:, 1:] |= bad[:, :-1]
        touch[:, :-1] |= bad[:, 1:]
        keep = ~touch[keep_rows]
        fto = r["ftle"][keep_rows]
        den = np.linalg.norm(fto[keep])
        res.update({
            "rows": [int(a), int(b)], "straddles_rank_cut": world > 1,
            "sample_rows_bit_identical_to_full_grid_launch": identical,
            "ftle_rel_l2": float(np.linalg.norm((ft_gpu - fto)[keep]) / den) if den > 0 else 0.0,
            "ftle_pixels_compared": int(keep.sum()), "ftle_pixels_excluded": int((~keep).sum()),
            "ftle_max_abs_diff": float(np.abs(ft_gpu - fto)[keep].max()),
            "gates": {"max_rel_dx_matching": 1e-8, "ftle_rel_l2": 1e-6, "mismatch_fraction": 1e-5},
            "oracle": "oracle/ C restatement, glibc libm, unfused arithmetic, same rtol/atol",
        })
        res["pass"] = bool(identical and res["max_rel_dx_matching"] <= 1e-8 and res["ftle_rel_l2"] <= 1e-6
                           and res["mismatch_fraction"] <= 1e-5)
        parity = res
        if world == 1:
            cpu = {"value": r["pps"], "unit": "grid points/s", "cores": r["threads"], "kind": "port",
                   "sample": f"{rows_s} contiguous rows x {n} columns of the same grid "
                             f"({rows_s * n} particles, best of 3: {r['seconds']:.1f} s), oracle/ C port with OpenMP",
                   "host_cpus": os.cpu_count()}
    if world > 1:
        dist.barrier()
    return parity, cpu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=16384, help="grid is n x n (BASELINE: 16384)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ridges", action="store_true", help="skip the ridge-extraction tail (N = 1)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
