// flowmap_kernel.cuh -- K1 / K1n: the particle flow-map kernel template (one thread per particle).
//
// Replaces the prange loops of /root/reference/src/numbacs/integration.py:
//   flowmap 46-52, flowmap_n 105-111, flowmap_grid_2D 166-172, flowmap_n_grid_2D 517-523.
// Thread q owns particle q of the C-order 'ij' grid (j fastest), so a warp is 32 consecutive
// y-neighbours: their trajectories, step counts and (for the spline flow) coefficient cells are
// similar, and the final 16-byte stores are fully coalesced.
// One translation unit per flow kind instantiates it (flowmap_<kind>.cu) so they build in parallel.
#pragma once
#include <atomic>

#include "common.cuh"
#include "dop853.cuh"
#include "flows.cuh"
#include "launch.cuh"

namespace b200cs {

namespace {

// Launch shape per (flow, output mode).  Default: 128-thread blocks, as many as the registers allow.
// The double-gyre kernels (the headline workload) pin that to five blocks per SM (96 registers): with
// the guarded sines six blocks at 80 registers were 0.5 % ahead (profiles/r2_ab_dg.txt), with the
// branch-free sines the larger basic blocks want the registers (1103 vs 1082 M points/s at 8192^2,
// profiles/r2_ab_dg_nobranch.txt).  Their attempt loop fits the 32 KB L1.5 instruction cache, so free-running small
// blocks beat every lockstep shape (profiles/r1c_ab_variants_a.txt: 640-thread lockstep 879,
// 2 x 384 lockstep 907, free 128 x 5 / 192 x 4 / 256 x 3 / 64 x 12 all 915-917 M points/s at 8192^2).
// The spline kernels gain 32 % from LOCKSTEP -- one 512-thread block per SM whose warps meet at a
// barrier every 8 attempts and so share the instruction-cache footprint of a loop that is several
// times the cache (their 64-tap RHS is unrolled into every stage).
template <class Rhs, bool DENSE>
struct KernelShape {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 1;
    static constexpr bool kLockstep = false;
    static constexpr bool kSlopesInSmem = false;
};
// Lane -> particle mapping of the final-time grid kernels.  kTileI == 1: a warp is 32 consecutive j
// (one grid row segment).  kTileI == 4 / 8: a warp is a kTileI x (32 / kTileI) tile of the grid.  A
// warp runs for as long as its slowest lane, and the attempt count of a particle is correlated
// with its neighbours' in BOTH directions, so a compact tile wastes fewer lane-attempts than a
// 32 x 1 strip (oracle attempt counts of config 2, bickley 2001 x 601, T = 6: lane efficiency 0.81
// for 1 x 32, 0.87 for 4 x 8 / 8 x 4); the stores stay whole 128-byte lines for kTileI <= 4.
// Measured (profiles/r3_ab_tile_hinit.txt; 1 x 32 -> 4 x 8 -> 8 x 4): double gyre 8192^2 1110.8 ->
// 1126.8 -> 1124.2, Bickley config 2 250.1 -> 271.3 -> 275.4 (289.6 -> 306.3 -> 311.2 at 10.8 M
// particles), spline probe 373.8 -> 389.2 -> 382.8 M points/s; outputs bit-identical.
#ifndef B200CS_TILE_I
#define B200CS_TILE_I 4
#endif
#ifndef B200CS_BICKLEY_TILE_I
#define B200CS_BICKLEY_TILE_I 8
#endif
#ifndef B200CS_DG_TILE_I
#define B200CS_DG_TILE_I B200CS_TILE_I
#endif
#ifndef B200CS_SPLINE_TILE_I
#define B200CS_SPLINE_TILE_I B200CS_TILE_I
#endif
// kQueue (optional member of a KernelShape): the final-time grid / point-list launches of this flow
// run as the QUEUE kernels below (lane-level work fetch) instead of one particle per thread.
#ifndef B200CS_BICKLEY_QUEUE
#define B200CS_BICKLEY_QUEUE 1
#endif
#ifndef B200CS_DG_QUEUE
#define B200CS_DG_QUEUE 0
#endif
#ifndef B200CS_SPLINE_QUEUE
#define B200CS_SPLINE_QUEUE 0
#endif
// Launch shape of the Bickley queue kernel: ONE 640-thread lockstep block per SM (96 registers).  The
// unrolled attempt of this flow is 49 KB of code against a 32 KB instruction cache (ncu on the free
// 128-thread shape: no_instruction 2.2 cycles per issue, the largest stall after the fixed-latency
// wait); in lockstep the 20 warps of an SM stream through it together.  What made lockstep lose in
// round 2 -- warps that ran out of particles idling at the barrier until the block's slowest lane
// was done -- is gone with the queue: every warp refills until the whole launch is drained.
// Measured on config 2 / at 10.8 M particles (profiles/r3_ab_lockstep_queue.txt): free 128 x 5
// 288.9 / 319.1, lockstep 640 x 1 re-aligned every 8 / 2 / 1 attempts 317.9 / 319.8 / 314.9 and
// 351.1 / 360.6 / 354.1, 512 x 1 (128 registers) 318.3 / 338.6, 320 x 2 287.6 / 321.5 M points/s.
#ifndef B200CS_BICKLEY_QTHREADS
#define B200CS_BICKLEY_QTHREADS 640
#endif
#ifndef B200CS_BICKLEY_QMINBLOCKS
#define B200CS_BICKLEY_QMINBLOCKS 1
#endif
#ifndef B200CS_BICKLEY_QLOCKSTEP
#define B200CS_BICKLEY_QLOCKSTEP true
#endif
template <class T, class = void>
struct shape_queue : std::false_type {};
// launch shape of the queue kernel: kQThreads / kQMinBlocks / kQLockstep of the KernelShape, else its plain shape
template <class T, class = void>
struct queue_shape {
    static constexpr int kThreads = T::kThreads, kMinBlocks = T::kMinBlocks;
    static constexpr bool kLockstep = false;
};
template <class T>
struct queue_shape<T, std::void_t<decltype(T::kQThreads)>> {
    static constexpr int kThreads = T::kQThreads, kMinBlocks = T::kQMinBlocks;
    static constexpr bool kLockstep = T::kQLockstep;
};
template <class T>
struct shape_queue<T, std::void_t<decltype(T::kQueue)>> : std::bool_constant<T::kQueue> {};
template <class T, class = void>
struct shape_tile_i : std::integral_constant<int, B200CS_TILE_I> {};
template <class T>
struct shape_tile_i<T, std::void_t<decltype(T::kTileI)>> : std::integral_constant<int, T::kTileI> {};
#ifndef B200CS_DG_THREADS
#define B200CS_DG_THREADS 128
#endif
#ifndef B200CS_DG_LOCKSTEP
#define B200CS_DG_LOCKSTEP false
#endif
#ifndef B200CS_DG_KSMEM   // A/B: stage slopes in shared memory to run 6 / 7 blocks per SM (80 / 72 registers):
#define B200CS_DG_KSMEM false   // 1044 / 1046 against 1150 M points/s at 8192^2 (profiles/r3_ab_dg_ksmem.txt)
#endif
#ifndef B200CS_DG_MINBLOCKS
#define B200CS_DG_MINBLOCKS 5
#endif
template <bool DAMPED>
struct KernelShape<DoubleGyreT<DAMPED>, false> {
    static constexpr int kThreads = B200CS_DG_THREADS;
    static constexpr int kMinBlocks = B200CS_DG_MINBLOCKS;
    static constexpr bool kLockstep = B200CS_DG_LOCKSTEP;
    static constexpr bool kSlopesInSmem = B200CS_DG_KSMEM;
    static constexpr int kTileI = B200CS_DG_TILE_I;
    static constexpr bool kQueue = B200CS_DG_QUEUE != 0;
};
// Bickley jet, final-time kernels: five blocks per SM (96 registers, 8 bytes of spills) measured
// 207.9 against 204.1 M points/s uncapped (128 registers, four blocks) on config 2
// (profiles/r1f_ab_bickley.txt)
#ifndef B200CS_BICKLEY_MINBLOCKS
#define B200CS_BICKLEY_MINBLOCKS 5
#endif
#ifndef B200CS_RHS_SLOTS   // out-of-line RHS reads its parameters from a __constant__ slot: 304 -> 331 M points/s (spline probe)
#define B200CS_RHS_SLOTS 1
#endif
template <class T, class = void>
struct rhs_out_of_line : std::false_type {};
template <class T>
struct rhs_out_of_line<T, std::void_t<decltype(T::kOutOfLine)>> : std::bool_constant<T::kOutOfLine> {};
#ifndef B200CS_BICKLEY_KSMEM
#define B200CS_BICKLEY_KSMEM false
#endif
#ifndef B200CS_SPLINE_KSMEM
#define B200CS_SPLINE_KSMEM false
#endif
#ifndef B200CS_BICKLEY_THREADS
#define B200CS_BICKLEY_THREADS 128
#endif
#ifndef B200CS_BICKLEY_LOCKSTEP
#define B200CS_BICKLEY_LOCKSTEP false
#endif
template <>
struct KernelShape<BickleyJet, false> {
    static constexpr int kThreads = B200CS_BICKLEY_THREADS;
    static constexpr int kMinBlocks = B200CS_BICKLEY_MINBLOCKS;
    static constexpr bool kLockstep = B200CS_BICKLEY_LOCKSTEP;
    static constexpr bool kSlopesInSmem = B200CS_BICKLEY_KSMEM;
    static constexpr int kTileI = B200CS_BICKLEY_TILE_I;
    static constexpr bool kQueue = B200CS_BICKLEY_QUEUE != 0;
    static constexpr int kQThreads = B200CS_BICKLEY_QTHREADS, kQMinBlocks = B200CS_BICKLEY_QMINBLOCKS;
    static constexpr bool kQLockstep = B200CS_BICKLEY_QLOCKSTEP;
};
// Round 2: with the 64-tap RHS out of line (B200CS_SPLINE_NOINLINE, flows.cuh) the attempt loop fits
// the instruction cache, so the spline kernels run as free 128-thread blocks, five per SM (96
// registers), like the analytic flows: 287 -> 304 M points/s on the 2701 x 1001 config-3 probe
// (profiles/r2_ab_spline.txt; lockstep 512 x 1 was the round-1 shape, kept as an A/B option).
// What bounds the kernel now is the L1 data pipe: 64 taps x 16 B per lane and RHS = 256 wavefronts
// per warp-RHS, measured at 60 % of the pipe's peak (profiles/r2a_flowmap_spline_c3.txt); parking
// the stage slopes in shared memory (B200CS_SPLINE_KSMEM) to free registers changed nothing.
#ifndef B200CS_SPLINE_THREADS
#define B200CS_SPLINE_THREADS 128
#endif
#ifndef B200CS_SPLINE_MINBLOCKS
#define B200CS_SPLINE_MINBLOCKS 5
#endif
#ifndef B200CS_SPLINE_LOCKSTEP
#define B200CS_SPLINE_LOCKSTEP false
#endif
template <int SPH>
struct KernelShape<Spline2D<SPH, false>, false> {
    static constexpr int kThreads = B200CS_SPLINE_THREADS;   // 512 x 128 registers = the whole register file
                                                             // (measured: 448 -> 246, 512 -> 281, 640 (96 regs, spills) -> 276 M points/s)
    static constexpr int kMinBlocks = B200CS_SPLINE_MINBLOCKS;
    static constexpr bool kLockstep = B200CS_SPLINE_LOCKSTEP;
    static constexpr bool kSlopesInSmem = B200CS_SPLINE_KSMEM;
    static constexpr int kTileI = B200CS_SPLINE_TILE_I;
    static constexpr bool kQueue = B200CS_SPLINE_QUEUE != 0;
};

template <int N>
struct RowSink {
    double *row;  // out + q*n*N
    bool vec;     // N == 2 and the base pointer is 16-byte aligned
    __device__ __forceinline__ void operator()(int k, const double (&v)[N]) const {
        if (N == 2 && vec) {
            *reinterpret_cast<double2 *>(row + 2 * (long long)k) = make_double2(v[0], v[1]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) row[(long long)k * N + i] = v[i];
        }
    }
};

// MODE: kModePts = point list pts[npts, N]; kModeGrid = 'ij' grid (x[i], y[j]);
// kModeAux = auxiliary stencil of flowmap_aux_grid_2D (integration.py:249-464): particle
// q = (i*ny + j)*n_aux + k starts at (x[i], y[j]) + aux_grid[k], aux_grid = [(h,0), (-h,0), (0,h),
// (0,-h), (0,0)]; the n_aux stencil points of a cell are neighbouring lanes (their trajectories
// and step sequences are practically identical, so they do not diverge).
// kModeSeries = the 'ij' grid integrated from nt different initial times in ONE launch (FTLE time
// series, the intermediate maps of flowmap_composition_initial): particle q belongs to frame
// q / (nx*ny) and integrates over [t0s[frame], t0s[frame] + T].  A 201 x 101 movie frame is only
// 20 k particles -- a seventh of one wave of the machine -- so batching the frames is what fills it.
constexpr int kModePts = 0, kModeGrid = 1, kModeAux = 2, kModeSeries = 3;

template <class Rhs, bool DENSE, int MODE>
struct grid_tile_i {
    static constexpr int value =
        (MODE == kModeGrid && !DENSE && Rhs::N == 2) ? shape_tile_i<KernelShape<Rhs, DENSE>>::value : 1;
};

template <class Rhs, bool DENSE, int MODE>
__global__ void __launch_bounds__(KernelShape<Rhs, DENSE>::kThreads, KernelShape<Rhs, DENSE>::kMinBlocks)
flowmap_kernel(const __grid_constant__ IntegArgs A) {
    constexpr int N = Rhs::N;
    constexpr int kBlock = KernelShape<Rhs, DENSE>::kThreads;
    constexpr bool kLockstep = KernelShape<Rhs, DENSE>::kLockstep;
    constexpr int kTI = grid_tile_i<Rhs, DENSE, MODE>::value;
    long long q = (long long)blockIdx.x * kBlock + threadIdx.x;
    bool in_range = q < A.npts;
    if constexpr (kTI > 1) {   // warp = kTI x (32 / kTI) tile of the grid (see shape_tile_i)
        constexpr int kTJ = 32 / kTI;
        const long long w = q >> 5;
        const int lane = threadIdx.x & 31;
        const long long tiles_j = (A.ny + kTJ - 1) / kTJ;
        const long long ti = w / tiles_j, tj = w - ti * tiles_j;
        const long long i = ti * kTI + lane / kTJ, j = tj * kTJ + lane % kTJ;
        in_range = i < A.nx && j < A.ny;
        q = in_range ? i * A.ny + j : 0;
    }
    bool active = in_range;
    if (MODE != kModeAux && MODE != kModeSeries && active && A.mask != nullptr) active = (A.mask[q] == 0);
    double x0 = A.x0, xend = A.xend;   // warp-uniform except in kModeSeries

    double y[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = 0.0;
    if (MODE == kModeAux) {
        if (active) {
            const long long cell = q / A.n_aux;
            const int k = (int)(q - cell * A.n_aux);
            const long long i = cell / A.ny, j = cell - i * A.ny;
            if (A.mask != nullptr) active = (A.mask[cell] == 0);
            // edge cells: nothing without compute_edge; with it all four points when there is no
            // centre point (n_aux == 4), else the centre point only (integration.py:320-343, 425-446)
            if (i == 0 || i == A.nx - 1 || j == 0 || j == A.ny - 1)
                active = active && A.aux_edge && (A.n_aux == 4 || k == 4);
            if (active) {
                const double ox = (k == 0) ? A.aux_h : (k == 1) ? -A.aux_h : 0.0;
                const double oy = (k == 2) ? A.aux_h : (k == 3) ? -A.aux_h : 0.0;
                y[0] = A.x[i] + ox;  // np.array([x[i], y[j]]) + aux_grid[k, :]
                y[1] = A.y[j] + oy;
            }
        }
    } else if (MODE == kModeSeries) {
        if (in_range) {
            const long long frame = q / A.frame_pts, ql = q - frame * A.frame_pts;
            if (A.mask != nullptr) active = (A.mask[ql] == 0);
            const double t0 = A.t0s[frame], p0 = A.rhs.p[0];
            x0 = p0 * t0;                // params[0] * linspace(t0, t0 + T, 2)   (integration.py:164)
            xend = p0 * (t0 + A.series_T);
            if (active) {
                const long long i = ql / A.ny, j = ql - i * A.ny;
                y[0] = A.x[i];
                y[1] = A.y[j];
            }
        }
    } else if (active) {
        if (MODE == kModeGrid) {
            const long long i = q / A.ny, j = q - i * A.ny;
            y[0] = A.x[i];
            y[1] = A.y[j];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) y[i] = A.pts[q * N + i];
        }
    }

    StepCounts cnt;
    int status = B200CS_ST_MASKED;
    const long long row_len = DENSE ? (long long)A.n_out * N : N;
    double *row = A.out + q * row_len;

    const Rhs rhs(A.rhs);
    const bool integrate = active && (xend != x0);   // T == 0: the flow map is the identity
    constexpr bool kKSmem = KernelShape<Rhs, DENSE>::kSlopesInSmem;
    __shared__ double slopes_smem[kKSmem ? (DENSE ? 17 : 14) * N * kBlock : 1];
    auto slopes = [&] {
        if constexpr (kKSmem) return SmemSlopes<N>{slopes_smem + threadIdx.x, kBlock};
        else return RegSlopes<N>{};
    };
    if (DENSE) {
        RowSink<N> sink{row, A.out_aligned16 != 0};
        if (active) {
            sink(0, y);  // row 0 is the initial condition
            if (!integrate)
                for (int k = 1; k < A.n_out; ++k) sink(k, y);
        } else if (in_range) {
            for (long long k = 0; k < row_len; ++k) row[k] = 0.0;  // masked: zeros (integration.py:163, 515)
        }
        status = dop853_integrate<true, kLockstep>(rhs, integrate, y, x0, xend, A.rtol, A.atol, A.n_out,
                                                   A.out_p0, A.out_t0, A.out_step, sink, cnt, slopes());
    } else {
        status = dop853_integrate<false, kLockstep>(rhs, integrate, y, x0, xend, A.rtol, A.atol, 0, 0.0, 0.0,
                                                    0.0, NoSink<N>{}, cnt, slopes());
    }
    if (active && !integrate) status = B200CS_ST_OK;

    if (in_range) {
        if (!DENSE) {
            if (N == 2 && A.out_aligned16) {
                *reinterpret_cast<double2 *>(row) = make_double2(y[0], y[1]);
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) row[i] = y[i];
            }
        }
        if (A.status) A.status[q] = status;
        if (A.steps) {
            A.steps[2 * q] = cnt.accepted;
            A.steps[2 * q + 1] = cnt.rejected;
        }
    }
    if (A.stats) {
        const int nstep = cnt.accepted + cnt.rejected;
        unsigned long long nfev =
            integrate ? 2ull + 11ull * nstep + cnt.accepted + 3ull * cnt.dense : 0ull;
        unsigned long long acc = cnt.accepted, rej = cnt.rejected;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nfev += __shfl_down_sync(0xffffffffu, nfev, o);
            acc += __shfl_down_sync(0xffffffffu, acc, o);
            rej += __shfl_down_sync(0xffffffffu, rej, o);
        }
        if ((threadIdx.x & 31) == 0 && nfev) {
            atomicAdd(&A.stats[0], nfev);
            atomicAdd(&A.stats[1], acc);
            atomicAdd(&A.stats[2], rej);
        }
    }
}

// ---- K1n + K3 fused: LAVD carried along the trajectory ------------------------------------------
// Replaces flowmap_n_grid_2D followed by lavd_grid_2D's particle loop
// (integration.py:467-533, diagnostics.py:336-377; Simpson rule utils.py:611-655) without ever
// materialising the [nx, ny, n, 2] trajectory array (10 GB at config 4): the dense-output sink
// evaluates |vort(t_k, traj(t_k)) - vort_avg[k]| as row k is produced and accumulates the composite
// Simpson sum on the fly (the three last integrand values are kept for the odd-interval rule).
// VK = 1: the vorticity is a cubic spline with time-collapsed slabs (what get_callable_scalar fields
// always are when the slabs fit): the evaluator is called directly, the trilinear / 3-D variants and
// their run-time switches are not in the loop.  VK = 0: anything (scalar_at_k decides per call).
template <int VK>
struct LavdSink {
    const ScalarDev *S;
    const double *tspan, *vavg;
    double px, py;
    int n;             // number of output times
    int m_simpson;     // number of intervals covered by the 1/3 rule: n-1 if even, n-2 if odd
    double f0, sum, fl1, fl2, fl3;  // first value, weighted interior sum, last three values
    // |vort(t_k, position) - mean_k|  (diagnostics.py:350-376)
    __device__ __forceinline__ double integrand(int k, const double (&v)[2]) const {
        double x = v[0], y = v[1];
        if (px != 0.0) x = pymod_any(x, px);
        if (py != 0.0) y = pymod_any(y, py);
        if constexpr (VK == 1) return fabs(eval_spline_s2(S->g, S->W + (long long)k * S->wstride, x, y) - __ldg(vavg + k));
        else return fabs(scalar_at_k(*S, k, __ldg(tspan + k), x, y) - __ldg(vavg + k));
    }
    __device__ __forceinline__ void accumulate(int k, double f) {
        fl3 = fl2;
        fl2 = fl1;
        fl1 = f;
        if (k == 0) f0 = f;
        else if (k < m_simpson) sum = fma((k & 1) ? 4.0 : 2.0, f, sum);
    }
    __device__ __forceinline__ void operator()(int k, const double (&v)[2]) { accumulate(k, integrand(k, v)); }
    // composite Simpson (utils.py:611-655) with spacing h
    __device__ __forceinline__ double finish(double h) const {
        const int m = n - 1;
        if ((m & 1) == 0) return (f0 + fl1 + sum) * (h / 3.0);
        double val = (f0 + fl2 + sum) * (h / 3.0);
        val += (5.0 * h / 12.0) * fl1 + (2.0 * h / 3.0) * fl2 - (h / 12.0) * fl3;
        return val;
    }
};

// The fused LAVD kernel keeps 32 x 1 strips: its cost is the 64-tap VORTICITY gather at 601 output
// times per particle, whose lanes coalesce along y (config 4 before the time-collapsed slabs: 32.2 ms
// with strips, 36.4 ms with 4 x 8 tiles, profiles/r3_configs_c1_c4_before_slabs.json /
// r3_configs_c1_c4_lavd_tile4.json).
#ifndef B200CS_LAVD_TILE_I
#define B200CS_LAVD_TILE_I 1
#endif
// Four blocks per SM (128 registers; the spills sit in the ~2 step attempts of a particle, not in
// its 601 output evaluations): the kernel is bound by the latency of the per-output chain
// (dense polynomial -> cell location -> 16 gathers -> Simpson), config 4: 21.3 ms at 168 registers /
// three blocks, 14.6 at four, 15.6 at five; evaluating two output times per iteration for overlap
// (254 registers) measured 20.6 / 15.5 and was dropped (profiles/r3_ab_lavd.txt).
// A second instantiation of the fused LAVD kernel for cubic-spline vorticity on time-collapsed slabs
// (VK = 1 below): config 4 13.6 -> 13.1 ms (profiles/r3_ab_lavd.txt), results identical.
#ifndef B200CS_LAVD_SPECIALIZE
#define B200CS_LAVD_SPECIALIZE 1
#endif
#ifndef B200CS_LAVD_MINBLOCKS
#define B200CS_LAVD_MINBLOCKS 4
#endif
template <class Rhs, int VK>
__global__ void __launch_bounds__(128, B200CS_LAVD_MINBLOCKS) lavd_flowmap_kernel(const __grid_constant__ IntegArgs A) {
    static_assert(Rhs::N == 2, "LAVD is defined for 2-D flows");
    constexpr int kTI = B200CS_LAVD_TILE_I, kTJ = 32 / kTI;   // warp = kTI x kTJ tile of the grid (shape_tile_i)
    long long q = (long long)blockIdx.x * 128 + threadIdx.x;
    if (kTI > 1) {
        const long long w = q >> 5;
        const int lane = threadIdx.x & 31;
        const long long tiles_j = (A.ny + kTJ - 1) / kTJ;
        const long long ti = w / tiles_j, tj = w - ti * tiles_j;
        const long long i = ti * kTI + lane / kTJ, j = tj * kTJ + lane % kTJ;
        if (i >= A.nx || j >= A.ny) return;
        q = i * A.ny + j;
    }
    if (q >= A.npts) return;
    const bool active = (A.mask == nullptr || A.mask[q] == 0);
    double y[2] = {0.0, 0.0};
    StepCounts cnt;
    int status = B200CS_ST_MASKED;
    double lavd = 0.0;
    if (active) {
        const long long i = q / A.ny, j = q - i * A.ny;
        y[0] = A.x[i];
        y[1] = A.y[j];
        const int m = A.n_out - 1;
        LavdSink<VK> sink{&A.vort, A.tspan_phys, A.vort_avg, A.period_x, A.period_y, A.n_out,
                      (m & 1) ? m - 1 : m, 0.0, 0.0, 0.0, 0.0, 0.0};
        sink(0, y);
        const Rhs rhs(A.rhs);
        if (A.xend == A.x0) {
            status = B200CS_ST_OK;
            for (int k = 1; k < A.n_out; ++k) sink(k, y);
        } else {
            status = dop853_integrate<true, false>(rhs, true, y, A.x0, A.xend, A.rtol, A.atol, A.n_out,
                                                   A.out_p0, A.out_t0, A.out_step, sink, cnt, RegSlopes<Rhs::N>{});
        }
        lavd = sink.finish(fabs(A.tspan_phys[1] - A.tspan_phys[0]));
    }
    A.lavd[q] = lavd;
    if (A.out) {  // optional final positions
        A.out[2 * q] = y[0];
        A.out[2 * q + 1] = y[1];
    }
    if (A.status) A.status[q] = status;
    if (A.stats && active) {
        const int nstep = cnt.accepted + cnt.rejected;
        atomicAdd(&A.stats[0], 2ull + 11ull * nstep + cnt.accepted + 3ull * cnt.dense);
        atomicAdd(&A.stats[1], (unsigned long long)cnt.accepted);
        atomicAdd(&A.stats[2], (unsigned long long)cnt.rejected);
    }
}

template <class Rhs>
void launch_lavd_one(const IntegArgs &A, cudaStream_t s) {
    constexpr int kTI = B200CS_LAVD_TILE_I;
    long long threads = A.npts;
    if (kTI > 1) threads = ((A.nx + kTI - 1) / kTI) * ((A.ny + 32 / kTI - 1) / (32 / kTI)) * 32;
    const long long blocks = A.npts > 0 ? (threads + 127) / 128 : 0;
    if (blocks <= 0) return;
    B2_REQUIRE(blocks < 2147483647LL, "too many particles for one launch (%lld)", A.npts);
#if B200CS_LAVD_SPECIALIZE
    if (A.vort.W != nullptr && !A.vort.linear) lavd_flowmap_kernel<Rhs, 1><<<(unsigned)blocks, 128, 0, s>>>(A);
    else
#endif
    lavd_flowmap_kernel<Rhs, 0><<<(unsigned)blocks, 128, 0, s>>>(A);
    B2_CHECK_CUDA(cudaGetLastError());
}

// One __constant__ parameter slot per stream (c_rhs_slots, flows.cuh): launches on one stream are
// ordered, so re-writing the stream's slot before a launch cannot disturb an earlier kernel; two
// streams never share a slot until more than kRhsSlots distinct streams have been seen, at which
// point the device is drained once before the table is reused.
inline int rhs_slot_for_stream(cudaStream_t s) {
    static std::mutex mu;
    static std::vector<cudaStream_t> seen;
    std::lock_guard<std::mutex> lk(mu);
    for (size_t k = 0; k < seen.size(); ++k)
        if (seen[k] == s) return (int)k;
    if ((int)seen.size() == kRhsSlots) {
        cudaDeviceSynchronize();
        seen.clear();
    }
    seen.push_back(s);
    return (int)seen.size() - 1;
}

// ---- queue kernels: lane-level work fetch for flows whose particles need very different numbers
// of step attempts ---------------------------------------------------------------------------
// With one particle per thread a warp runs for as long as its slowest lane: on config 2 (Bickley jet
// 2001 x 601, T = 6; 6..61 attempts per particle) the oracle's attempt counts give a lane
// efficiency of 0.81 for 32 x 1 strips and 0.87 for 8 x 4 tiles, and ncu counts 25.2 of 32 lanes
// active.  Here the integration is split in two launches:
//   flowmap_init_kernel   one thread per particle SLOT (tile order): initial condition, mask, first
//                         slope and hinit -- warp-uniform code, nothing diverges -- parked as 48
//                         bytes of state per slot (SoA: y, first slope, h, output index; index -1
//                         marks a slot without work, whose zeros / MASKED status this kernel has
//                         already written);
//   flowmap_queue_kernel  a persistent grid; every lane that finishes its particle stores the result
//                         and takes the next slot from a global counter (one warp-aggregated
//                         atomicAdd per refill) at the top of the attempt loop, so only the tail of the
//                         whole launch idles lanes.
// Every particle is integrated by exactly the same instruction sequence as in flowmap_kernel: the
// results are bit-identical (tools/grid_hash.py, tests/test_gpu_parity.py::test_queue_kernel_*).
template <int TI, int MODE>
__device__ __forceinline__ long long slot_to_particle(long long qi, const IntegArgs &A) {
    if (MODE == kModeGrid) {
        constexpr int kTJ = 32 / TI;
        const long long w = qi >> 5;
        const int lane = (int)(qi & 31);
        const long long tiles_j = (A.ny + kTJ - 1) / kTJ;
        const long long ti = w / tiles_j, tj = w - ti * tiles_j;
        const long long i = ti * TI + lane / kTJ, j = tj * kTJ + lane % kTJ;
        return (i < A.nx && j < A.ny) ? i * A.ny + j : -1;
    }
    return qi < A.npts ? qi : -1;
}

template <class Rhs, int MODE>
struct queue_tile_i {
    static constexpr int raw = shape_tile_i<KernelShape<Rhs, false>>::value;
    static constexpr int value = (MODE == kModeGrid) ? (raw > 1 ? raw : 1) : 1;
};

template <class Rhs, int MODE>
__global__ void __launch_bounds__(128) flowmap_init_kernel(const __grid_constant__ IntegArgs A) {
    static_assert(Rhs::N == 2, "queue kernels: 2-D flows");
    constexpr int kTI = queue_tile_i<Rhs, MODE>::value;
    const long long qi = (long long)blockIdx.x * 128 + threadIdx.x;
    if (qi >= A.nq) return;
    const long long q = slot_to_particle<kTI, MODE>(qi, A);
    bool active = q >= 0;
    if (active && A.mask != nullptr) active = (A.mask[q] == 0);
    double y[2] = {0.0, 0.0}, k1[2] = {0.0, 0.0}, h = 0.0;
    if (active) {
        if (MODE == kModeGrid) {
            const long long i = q / A.ny, j = q - i * A.ny;
            y[0] = A.x[i];
            y[1] = A.y[j];
        } else {
            y[0] = A.pts[2 * q];
            y[1] = A.pts[2 * q + 1];
        }
        const Rhs rhs(A.rhs);
        RegSlopes<2> K;
        const double posneg = (A.xend - A.x0) < 0.0 ? -1.0 : 1.0;
        h = dop853_start(rhs, A.x0, y, A.rtol, A.atol, fabs(A.xend - A.x0), posneg, K);
        k1[0] = K[1][0];
        k1[1] = K[1][1];
    } else if (q >= 0) {   // masked: zeros (integration.py:163), final here
        if (A.out_aligned16) *reinterpret_cast<double2 *>(A.out + 2 * q) = make_double2(0.0, 0.0);
        else A.out[2 * q] = A.out[2 * q + 1] = 0.0;
        if (A.status) A.status[q] = B200CS_ST_MASKED;
        if (A.steps) A.steps[2 * q] = A.steps[2 * q + 1] = 0;
    }
    A.qstate[qi] = y[0];
    A.qstate[A.nq + qi] = y[1];
    A.qstate[2 * A.nq + qi] = k1[0];
    A.qstate[3 * A.nq + qi] = k1[1];
    A.qstate[4 * A.nq + qi] = h;
    // slot -> particle (output index), -1 for a slot without work: the refill needs no index arithmetic
    reinterpret_cast<long long *>(A.qstate)[5 * A.nq + qi] = active ? q : -1;
}

template <int TI, int MODE>
struct QueueFeeder {
    static constexpr bool kActive = true;
    const IntegArgs &A;
    long long q = -1;     // particle in flight (output index), -1: none
    bool done = false;    // the counter has passed the last slot (warp-uniform)
    unsigned long long nfev = 0, acc = 0, rej = 0;
    __device__ __forceinline__ explicit QueueFeeder(const IntegArgs &A_) : A(A_) {}
    __device__ __forceinline__ bool drained() const { return done; }
    // Called by ALL lanes of the warp.  A lane that is not alive first hands in the particle it holds
    // (final state, status, step counts), then takes the next slot with work, if any: returns true
    // with y, h and the first slope k1 of the new particle.
    __device__ __forceinline__ bool refill(bool alive, double (&y)[2], double &h, double (&k1)[2], int status,
                                           const StepCounts &cnt) {
        if (!alive && q >= 0) {
            if (A.out_aligned16) *reinterpret_cast<double2 *>(A.out + 2 * q) = make_double2(y[0], y[1]);
            else { A.out[2 * q] = y[0]; A.out[2 * q + 1] = y[1]; }
            if (A.status) A.status[q] = status;
            if (A.steps) {
                A.steps[2 * q] = cnt.accepted;
                A.steps[2 * q + 1] = cnt.rejected;
            }
            nfev += 2ull + 11ull * (unsigned)(cnt.accepted + cnt.rejected) + (unsigned)cnt.accepted;
            acc += (unsigned)cnt.accepted;
            rej += (unsigned)cnt.rejected;
            q = -1;
        }
        if (done) return false;
        const unsigned lane = threadIdx.x & 31;
        const unsigned idle = __ballot_sync(0xffffffffu, !alive);
        const int n = __popc(idle);
        const int leader = __ffs(idle) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(A.qcounter, (unsigned long long)n);
        base = __shfl_sync(0xffffffffu, base, leader);
        if ((long long)(base + n) >= A.nq) done = true;
        if (alive) return false;
        const long long qi = (long long)base + __popc(idle & ((1u << lane) - 1u));
        if (qi >= A.nq) return false;
        const long long qn = reinterpret_cast<const long long *>(A.qstate)[5 * A.nq + qi];
        if (qn < 0) return false;   // masked, or an empty slot of an edge tile
        y[0] = A.qstate[qi];
        y[1] = A.qstate[A.nq + qi];
        k1[0] = A.qstate[2 * A.nq + qi];
        k1[1] = A.qstate[3 * A.nq + qi];
        h = A.qstate[4 * A.nq + qi];
        q = qn;
        return true;
    }
};

template <class Rhs, int MODE>
__global__ void __launch_bounds__(queue_shape<KernelShape<Rhs, false>>::kThreads,
                                  queue_shape<KernelShape<Rhs, false>>::kMinBlocks)
flowmap_queue_kernel(const __grid_constant__ IntegArgs A) {
    constexpr int kTI = queue_tile_i<Rhs, MODE>::value;
    constexpr bool kLockstep = queue_shape<KernelShape<Rhs, false>>::kLockstep;
    const Rhs rhs(A.rhs);
    QueueFeeder<kTI, MODE> feeder(A);
    double y[2] = {0.0, 0.0};
    StepCounts cnt;
    dop853_integrate<false, kLockstep>(rhs, true, y, A.x0, A.xend, A.rtol, A.atol, 0, 0.0, 0.0, 0.0, NoSink<2>{}, cnt,
                                       RegSlopes<2>{}, feeder);
    if (A.stats) {
        unsigned long long nfev = feeder.nfev, acc = feeder.acc, rej = feeder.rej;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nfev += __shfl_down_sync(0xffffffffu, nfev, o);
            acc += __shfl_down_sync(0xffffffffu, acc, o);
            rej += __shfl_down_sync(0xffffffffu, rej, o);
        }
        if ((threadIdx.x & 31) == 0 && nfev) {
            atomicAdd(&A.stats[0], nfev);
            atomicAdd(&A.stats[1], acc);
            atomicAdd(&A.stats[2], rej);
        }
    }
}

// smallest launch that goes through the queue kernels (below it one wave of flowmap_kernel is as good)
constexpr long long kQueueMinParticles = 1 << 16;

template <class Rhs, int MODE>
void launch_queue(const IntegArgs &A0, cudaStream_t s) {
    constexpr int kTI = queue_tile_i<Rhs, MODE>::value;
    constexpr int kBlock = queue_shape<KernelShape<Rhs, false>>::kThreads;
    IntegArgs A = A0;
    A.nq = A.npts;
    if (MODE == kModeGrid) A.nq = ((A.nx + kTI - 1) / kTI) * ((A.ny + 32 / kTI - 1) / (32 / kTI)) * 32;
    Scratch st((size_t)A.nq * 6 * sizeof(double) + 64, s);
    A.qstate = static_cast<double *>(st.ptr);
    A.qcounter = reinterpret_cast<unsigned long long *>(A.qstate + 6 * A.nq);
    B2_CHECK_CUDA(cudaMemsetAsync(A.qcounter, 0, sizeof(unsigned long long), s));
    const long long init_blocks = (A.nq + 127) / 128;
    B2_REQUIRE(init_blocks < 2147483647LL, "too many particles for one launch (%lld)", A.npts);
    // blocks of the queue kernel the current device holds at once (cached per device)
    static std::atomic<int> resident_of[64];
    int dev = 0;
    B2_CHECK_CUDA(cudaGetDevice(&dev));
    int resident = (dev >= 0 && dev < 64) ? resident_of[dev].load(std::memory_order_relaxed) : 0;
    if (resident == 0) {
        int sms = 0, per_sm = 0;
        B2_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        B2_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flowmap_queue_kernel<Rhs, MODE>, kBlock, 0));
        resident = sms * (per_sm > 0 ? per_sm : 1);
        if (dev >= 0 && dev < 64) resident_of[dev].store(resident, std::memory_order_relaxed);
    }
    const long long want = (A.nq + kBlock - 1) / kBlock;
    const unsigned blocks = (unsigned)(want < resident ? want : resident);
#if B200CS_RHS_SLOTS
    if constexpr (rhs_out_of_line<Rhs>::value) {
        A.rhs.slot = rhs_slot_for_stream(s);
        B2_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_rhs_slots, &A.rhs, sizeof(RhsParams), sizeof(RhsParams) * A.rhs.slot,
                                              cudaMemcpyHostToDevice, s));
    }
#endif
    flowmap_init_kernel<Rhs, MODE><<<(unsigned)init_blocks, 128, 0, s>>>(A);
    B2_CHECK_CUDA(cudaGetLastError());
    flowmap_queue_kernel<Rhs, MODE><<<blocks, kBlock, 0, s>>>(A);
    B2_CHECK_CUDA(cudaGetLastError());
}

template <class Rhs, bool DENSE, int MODE>
void launch_one(const IntegArgs &A, cudaStream_t s) {
    constexpr int kBlock = KernelShape<Rhs, DENSE>::kThreads;
    constexpr int kTI = grid_tile_i<Rhs, DENSE, MODE>::value;
    long long threads = A.npts;
    if constexpr (kTI > 1) threads = ((A.nx + kTI - 1) / kTI) * ((A.ny + 32 / kTI - 1) / (32 / kTI)) * 32;
    const long long blocks = A.npts > 0 ? (threads + kBlock - 1) / kBlock : 0;
    if (blocks <= 0) return;
    if constexpr (shape_queue<KernelShape<Rhs, DENSE>>::value && !DENSE && Rhs::N == 2 &&
                  (MODE == kModeGrid || MODE == kModePts)) {
        if (A.npts >= kQueueMinParticles && A.xend != A.x0 && A.out != nullptr) {
            launch_queue<Rhs, MODE>(A, s);
            return;
        }
    }
    B2_REQUIRE(blocks < 2147483647LL, "too many particles for one launch (%lld)", A.npts);
#if B200CS_RHS_SLOTS
    if constexpr (rhs_out_of_line<Rhs>::value) {   // the flows with an out-of-line RHS (spline / linear)
        IntegArgs B = A;
        B.rhs.slot = rhs_slot_for_stream(s);
        B2_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_rhs_slots, &B.rhs, sizeof(RhsParams), sizeof(RhsParams) * B.rhs.slot,
                                              cudaMemcpyHostToDevice, s));
        flowmap_kernel<Rhs, DENSE, MODE><<<(unsigned)blocks, kBlock, 0, s>>>(B);
        B2_CHECK_CUDA(cudaGetLastError());
        return;
    }
#endif
    flowmap_kernel<Rhs, DENSE, MODE><<<(unsigned)blocks, kBlock, 0, s>>>(A);
    B2_CHECK_CUDA(cudaGetLastError());
}

template <class Rhs>
void launch_rhs(const IntegArgs &A, int mode, cudaStream_t s) {
    const bool dense = A.n_out >= 2;
    if (mode == kModeAux || mode == kModeSeries) {
        if constexpr (Rhs::N == 2) {
            B2_REQUIRE(!dense, "the aux-grid / time-series flow maps are final-time quantities");
            if (mode == kModeAux) launch_one<Rhs, false, kModeAux>(A, s);
            else launch_one<Rhs, false, kModeSeries>(A, s);
        } else {
            B2_REQUIRE(false, "the aux grid and the time series need a 2-D flow");
        }
    } else if (dense) {
        if (mode == kModeGrid) launch_one<Rhs, true, kModeGrid>(A, s);
        else launch_one<Rhs, true, kModePts>(A, s);
    } else {
        if (mode == kModeGrid) launch_one<Rhs, false, kModeGrid>(A, s);
        else launch_one<Rhs, false, kModePts>(A, s);
    }
}

}  // namespace

}  // namespace b200cs
