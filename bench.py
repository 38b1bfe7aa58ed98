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
          pinned memory every step, the FTLE block copied device->host into pinned memory.
  roofline : the integration kernel (FP64-FMA bound; neither HBM nor tensor), ALGORITHMIC flops
          F = N_att*346 + N_fev*35 + 40 per particle from the kernel's own step
          statistics, divided by the kernel time measured live with CUDA events; peak = FP64 DFMA
          peak measured in this run (MEASURED_PEAKS.json has no FP64 entry).
  roofline_ftle : the FTLE stencil kernel against the measured HBM copy bandwidth (24 B/pixel).
  ridge_tail    : config 5's "+ FTLE ridge extraction" (C_eig_2D -> ftle_from_eig -> ftle_ridge_pts on
          the step's flow map), timed on its own after the K steps and reported beside the metric.
  cpu_baseline  : the CPU oracle (a port of the reference algorithm, oracle/) on all host threads
          over a bounded contiguous row sample of the same grid.

--impl reference times the reference's CPU implementation of the path.  The reference is pure
Python + numba whose solver / spline live in third-party packages that are not installed and
cannot be installed offline, so the arm runs the oracle port (kind "port") on all host cores.
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

def cpu_sample(n, rows, reps=1):
    """Oracle flow map + FTLE on `rows` contiguous rows (plus one halo row each side) of the
    n x n grid, all host threads.  Returns (points per second, seconds per rep, threads)."""
    import oracle as O
    O.build()
    f, p, _ = O.get_predefined_flow("double_gyre", int_direction=-1.0)
    x, y = np.linspace(0, 2, n), np.linspace(0, 1, n)
    i0 = n // 4
    xs = x[i0 - 1:i0 + rows + 1]
    best = float("inf")
    for _ in range(reps):
        t = time.perf_counter()
        fm = O.flowmap_grid_2D(f, T0, TINT, xs, y, p, rtol=RTOL, atol=ATOL)
        O.ftle_grid_2D(fm, TINT, x[1] - x[0], y[1] - y[0])
        best = min(best, time.perf_counter() - t)
    # the two halo rows are integrated too; count them as work done
    return (rows + 2) * n / best, best, O.num_threads()


def pick_rows(n, target_s):
    """Rows of the n x n grid the oracle integrates in about target_s seconds."""
    pps, _, _ = cpu_sample(n, 6)
    rows = int(pps * target_s / n)
    return max(8, min(rows, n - 2))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # one CPU run per node; the other ranks exit without work
    n = args.n
    rows = args.cpu_rows or pick_rows(n, 6.0)
    for _ in range(args.warmup):
        cpu_sample(n, max(8, rows // 8))
    times = []
    for _ in range(args.steps):
        _, t, threads = cpu_sample(n, rows)
        times.append(t)
    t_step = float(np.mean(times))
    val = (rows + 2) * n / t_step
    sample = f"{rows + 2} contiguous rows x {n} columns of the {n}x{n} grid per step ({(rows + 2) * n} particles)"
    line = {
        "impl": "reference", "metric": "FTLE grid points/s (flowmap+FTLE)", "value": val,
        "unit": "grid points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "sample": sample},
        "cpu_baseline": {"value": val, "unit": "grid points/s", "cores": threads, "kind": "port",
                         "sample": sample,
                         "note": "oracle/ C restatement of the reference algorithm (OpenMP over particles); "
                                 "the reference's own numba path cannot run: numbalsoda / interpolation "
                                 "are not installed and there is no network"},
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

def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from numbacs_b200 import _build, _lib
    from numbacs_b200.flows import get_predefined_flow
    from numbacs_b200.sharded import balanced_row_blocks, estimate_row_cost, exchange_halo_rows

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
    f, params, _ = get_predefined_flow("double_gyre", int_direction=-1.0)
    # row blocks of equal estimated COST (step attempts), not equal size: a 256 x 256 subsample of
    # the grid is integrated once (planning, outside the timed steps; every rank gets the same cuts)
    t_plan = time.perf_counter()
    if world > 1:
        cost = estimate_row_cost(f, T0, TINT, np.linspace(0, 2, n), np.linspace(0, 1, n), params, RTOL, ATOL)
        blocks = balanced_row_blocks(cost, world)
    else:
        blocks = [(0, n)]
    t_plan = time.perf_counter() - t_plan
    i0, i1 = blocks[rank]
    rows = i1 - i0
    has_lo, has_hi = int(rank > 0), int(rank < world - 1)

    x_host = torch.linspace(0, 2, n, dtype=torch.float64).pin_memory()
    y_host = torch.linspace(0, 1, n, dtype=torch.float64).pin_memory()
    x_dev, y_dev = x_host.cuda(), y_host.cuda()
    x_rows_dev = x_dev[i0:i1].contiguous()
    slab = torch.empty((has_lo + rows + has_hi, n, 2), dtype=torch.float64, device="cuda")
    own = slab[has_lo:has_lo + rows]
    ftle = torch.empty((rows, n), dtype=torch.float64, device="cuda")
    ftle_host = torch.empty((rows, n), dtype=torch.float64).pin_memory()
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
        _lib.check(L.b200cs_ftle_slab_2d(C.c_void_p(slab.data_ptr()), slab.shape[0], n, TINT, dx, dy,
                                         None, has_lo, has_hi, C.c_void_p(ftle.data_ptr()), sptr))

    def step(xr, yv, with_stats=False):
        k_flowmap(xr, yv, with_stats)
        if world > 1:
            exchange_halo_rows(slab, has_lo, has_hi, rank)
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
        if world > 1:
            exchange_halo_rows(slab, has_lo, has_hi, rank)
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

    # ---- end-to-end region: ONE C-ABI call per step with HOST buffers only -- x / y slab in from
    # pinned memory, the FTLE block out to pinned memory (b200cs_flowmap_ftle_grid_2d integrates in
    # row chunks and streams finished FTLE rows over PCIe while the next chunk is integrated).
    # Multi-GPU: every rank passes its row block plus one stencil-only halo row per interior edge
    # (recomputed, 2 of nx/N rows), so the end-to-end path needs no exchange at all.
    xs_host = x_host[i0 - has_lo:i1 + has_hi]
    def e2e_step():
        _lib.check(L.b200cs_flowmap_ftle_grid_2d(
            f, T0, TINT, C.c_void_p(xs_host.data_ptr()), xs_host.shape[0], C.c_void_p(y_host.data_ptr()), n,
            C.c_void_p(p_arr.ctypes.data), len(p_arr), 0, RTOL, ATOL, None, dx, dy, has_lo, has_hi,
            None, C.c_void_p(ftle_host.data_ptr()), None, None, sptr))

    e2e_step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        e2e_step()
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    checksum = float(ftle_host.sum()) if rows else 0.0   # device->host result actually read

    # ---- config 5's tail, reported beside the metric (not inside it): Cauchy-Green eigen-pairs ->
    # FTLE from the largest eigenvalue -> sub-pixel ridge points, on the flow map of the last step
    # (examples/ftle/plot_dg_ftle_ridges.py:43-72: percentile 0, sdd_thresh 10).  N = 1 only.
    ridge = None
    if world == 1 and not args.no_ridges:
        from numbacs_b200.diagnostics import C_eig_2D, ftle_from_eig
        from numbacs_b200.extraction import ftle_ridge_pts
        fm_full = slab[has_lo:has_lo + rows]

        def tail():
            vals, vecs = C_eig_2D(fm_full, dx, dy)
            ft2 = ftle_from_eig(vals[:, :, 1], TINT)
            return ft2, ftle_ridge_pts(ft2, vecs[:, :, :, 1], x_dev, y_dev, sdd_thresh=10.0, percentile=0)

        tail()
        best = float("inf")
        for _ in range(3):
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            ft2, rp = tail()
            r1.record(stream)
            torch.cuda.synchronize()
            best = min(best, r0.elapsed_time(r1))
        ridge = {"ms": best, "n_ridge_pts": int(rp.shape[0]),
                 "ftle_vs_ftle_grid_2D_rel_l2": float(torch.linalg.norm(ft2 - ftle) / torch.linalg.norm(ftle)),
                 "kernels": "C_eig_2D + ftle_from_eig + ftle_ridge_pts (detect, scan, compact), device-timed, "
                            "best of 3; sdd_thresh=10, percentile=0",
                 "algorithmic_bytes_per_point": 64 + 16 + 2 * 24}
        del ft2, rp
        torch.cuda.empty_cache()

    # ---- max over ranks
    t = torch.tensor([span_ms, e2e_ms, fm_ms, ft_ms, halo_ms], dtype=torch.float64, device="cuda")
    agg = torch.cat([torch.tensor(st, dtype=torch.float64, device="cuda"),
                     torch.tensor([checksum], dtype=torch.float64, device="cuda")])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    span_ms, e2e_ms, fm_ms, ft_ms, halo_ms = t.tolist()
    nfev, nacc, nrej, checksum = agg.tolist()

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
        # bounded CPU sample of the same workload on this box's host cores (rank 0, N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            rows_cpu = args.cpu_rows or pick_rows(n, 12.0)
            v, tsec, threads = cpu_sample(n, rows_cpu)
            cpu = {"value": v, "unit": "grid points/s", "cores": threads, "kind": "port",
                   "sample": f"{rows_cpu + 2} contiguous rows x {n} columns of the same grid "
                             f"({(rows_cpu + 2) * n} particles, {tsec:.1f} s), oracle/ C port with OpenMP",
                   "host_cpus": os.cpu_count()}
        line = {
            "metric": "FTLE grid points/s (flowmap+FTLE)", "value": value, "unit": "grid points/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "parallelism": f"row-block x{world}" + (", blocks balanced by estimated step count "
                                                              f"(planning {t_plan * 1e3:.1f} ms, not timed)" if world > 1 else ""),
                       "row_blocks": [list(b) for b in blocks],
                       "l2": "no flush: every step rewrites 24 B/point of outputs "
                             f"({24 * pts / world / 1e9:.2f} GB per GPU >> 126 MB L2), inputs are 2 x {n} doubles"},
            "e2e": {"value": e2e_val, "unit": "grid points/s",
                    "h2d_bytes_per_step": int(8 * (n + 2 * (world - 1) + n * world)),
                    "d2h_bytes_per_step": int(8 * pts), "ms_per_step": e2e_ms / K,
                    "result": "one b200cs_flowmap_ftle_grid_2d call per rank and step, host pointers only: x/y "
                              "from pinned memory, FTLE field to pinned memory (downloads overlap the "
                              "integration of later row chunks); the flow map stays in HBM"},
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
            "ridge_tail": ridge,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "halo_exchange_ms": halo_ms if world > 1 else 0.0,
            "wall_ms_per_step": t_wall / K * 1e3,
            "ftle_checksum": checksum,
            "fp64_peak_tflops_measured": fp64_peak,
        }
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
