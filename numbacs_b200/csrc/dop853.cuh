// dop853.cuh -- one-thread-per-particle adaptive Dormand-Prince 8(5,3) integrator.
//
// Replaces numbalsoda.dop853(funcptr, u0, t_eval, rtol, atol, data) as called from
// /root/reference/src/numbacs/integration.py:49, 108, 169, 520.  The step-size controller is
// Hairer's classical one (dop853.f): safe 0.9, fac1 0.333, fac2 6, beta 0, hinit with iord 8,
// "no growth after a reject", last step when x + 1.01 h passes xend; one continuous integration
// over the output times with 7th-order dense output (contd8) at interior times.  Parity with the
// reference requires the SAME accept/reject sequence per particle, so nothing in the controller
// is "improved": every particle keeps its own x, h and accept/reject history.
//
// GPU shape.  Every lane owns its particle, its step size and its decisions; one loop iteration
// is one step ATTEMPT (12 stages, error estimate, predicated state update), so a warp stays
// converged except for the FSAL evaluation of rejected lanes and the tail where lanes need
// different numbers of attempts.  The twelve stages are fully unrolled with the stage slopes in
// registers and the tableau's zero entries removed at compile time.
//
// What was measured on B200 before settling on this shape (profiles/README.md has the numbers):
// one warp-DFMA issues per two cycles per SM sub-partition (8.3-cycle dependent latency) and FP64
// instructions cannot take constant-bank operands on sm_100, so the kernel is bound by the FP64
// pipe AND by the issue slots its non-FP64 instructions take.  Variants with 2 or 4 particles per
// thread and the slopes in shared memory executed MORE instructions per particle and were
// 5-40 % slower; an out-of-line RHS cost more in calls than it saved in instruction fetch.
// What did pay, in order: a leaner RHS (own sine kernels, product-to-sum); batching the time-only
// part of the RHS over the stage times of a step; and -- for as long as the unrolled attempt loop
// was larger than the 32 KB L1.5 instruction cache -- LOCKSTEP blocks (one big block per SM that
// meets at a barrier every 8 attempts, so its warps share the loop's I-cache footprint).  Since
// round 1c the double-gyre loop is 1600 instructions = 25 KB (single-polynomial sines without
// parity select, constants folded on the host, damping compiled out): it fits the I-cache, so the
// double-gyre kernels run as free 128-thread blocks, five per SM (no barrier, no block tail where
// 8 % of the warp-time idled), FP64 pipe 59 % -> 82 % busy, 664 -> 1014 M points/s at 16384^2.
// Round 2: the spline kernels run free too (their 64-tap RHS went out of line); a warp of the
// final-time grid kernels is a 4 x 8 tile of the particle grid, not a 32 x 1 strip (fewer idle
// lane-attempts); and the Bickley jet, whose unrolled attempt is 49 KB, runs as a LOCKSTEP QUEUE
// kernel: one 640-thread block per SM whose lanes fetch the next pre-initialised particle when
// they finish (the Feeder hook below) -- lockstep shares the instruction-cache footprint, the
// queue removes what lockstep used to cost (flowmap_kernel.cuh).
#pragma once
#include <cuda_runtime.h>

#include <type_traits>
#include <utility>

#include "dop853_tableau.cuh"

// B200CS_STRICT: the parity-calibration build (tools/build_variant.py strict -DB200CS_STRICT=1
// -fmad=false, tests/test_gpu_parity_atsize.py).  Every expression of the integrator and of the
// right-hand sides is evaluated in the reference's (= the oracle's) order with separately rounded
// IEEE multiplications and additions, IEEE division / sqrt and CUDA libm sin / cos / cosh / tanh /
// pow.  What is left between this build and the CPU oracle is the difference between two libms
// (each <= 1-2 ulp): the distance no GPU implementation can go below, and the yardstick the
// product build's distance to the oracle is compared with.
#ifndef B200CS_STRICT
#define B200CS_STRICT 0
#endif
#ifndef B200CS_STRICT_INT   // the integrator half of the strict build on its own (A/B decomposition)
#define B200CS_STRICT_INT B200CS_STRICT
#endif
#if B200CS_STRICT_INT
#define B200CS_LEAN 0
#define B200CS_LEAN2 0
#endif
#ifndef B200CS_LEAN
#define B200CS_LEAN 1
#endif
#ifndef B200CS_LEAN2
#define B200CS_LEAN2 B200CS_LEAN
#endif
// B200CS_CTRL_FAST (round 2): the step-size controller without IEEE divisions -- 1/sk from the MUFU
// reciprocal seed + two Newton steps, and h_new = h * min(6, max(1/3, 0.9 err^(-1/8))) with
// err^(-1/8) from three MUFU rsqrt seeds + two Newton steps (~1e-16 relative), instead of
// pow_eighth + division by 0.9 + division of h by the factor.  Parity: the integrator-only /
// RHS-only decomposition (profiles/r2_parity_decomp.json) shows that ulp-level changes of the
// controller leave the distance to the oracle unchanged; the accept / reject test itself
// (err <= 1) is not touched.
#ifndef B200CS_CTRL_FAST
#define B200CS_CTRL_FAST (!B200CS_STRICT_INT)
#endif
// B200CS_CTRL_ONE_NEWTON (round 2, last session): one Newton step instead of two for 1/sk and
// err^(-1/8) (rcp_fast1 / inv_eighth_root1 below say why that cannot reach the parity figures):
// 804 -> 794 FP64 instructions per double-gyre attempt, 1132.9 -> 1148.8 M points/s at 8192^2,
// config-1 parity figures identical to the last digit (profiles/r3_ab_lockstep_queue.txt).
#ifndef B200CS_CTRL_ONE_NEWTON
#define B200CS_CTRL_ONE_NEWTON B200CS_CTRL_FAST
#endif
// lockstep QUEUE kernels re-align more often than the plain lockstep kernels: their warps never run
// dry, so a barrier costs only the skew of one or two attempts
#ifndef B200CS_QSYNC_EVERY
#define B200CS_QSYNC_EVERY 3   // every 2 / 3 / 4 attempts: 321.6 / 325.4 / 325.7 (config 2), 362.8 / 366.9 / 359.5 M points/s (10.8 M particles)
#endif
// B200CS_LEAN_TAIL: fewer issue slots in the accept / reject tail, results bit-identical
// (profiles/r3_ab_nonfp64.txt): 1 = the first-same-as-last slope is evaluated straight into K[1] in the
// final-time kernels and the last step leaves through the common exit (no early break: fewer
// merge copies), 245 -> 234 non-FP64 instructions, 1182.8 -> 1190.6 M points/s at 8192^2;
// 2 = also max(|y|, |y5|) as compare + select instead of fmax with its NaN fix-up: 223, 1193.7.
#ifndef B200CS_LEAN_TAIL
#define B200CS_LEAN_TAIL 2
#endif
// (Measured on top and dropped, profiles/r3_ab_nonfp64.txt: the step-size clamps as compare + select, +-0;
// the rsqrt of the error norm without libdevice's range test and slow-path call -- ptxas answers the
// merged basic block with 28 bytes of spills inside the loop: 1195 -> 1114 M points/s.)
// B200CS_DENSE_RCP: theta = (t_k - x) * (1/h) with ONE IEEE division per step instead of one per output
// row (contd8's s = (t - told)/h): a last-bit change of theta; config 4 (601 rows per particle)
// 14.2 -> 13.6 ms, float32 goldens of flowmap_n unchanged (profiles/r3_ab_lavd.txt).  Not in the
// strict build.
#ifndef B200CS_DENSE_RCP
#define B200CS_DENSE_RCP (!B200CS_STRICT_INT)
#endif
#ifndef B200CS_FSAL_ALWAYS
#define B200CS_FSAL_ALWAYS 0
#endif
// B200CS_HINIT_FAST (round 2, last session): hinit with the controller's reciprocal / rsqrt kernels
// instead of nine IEEE divisions and three IEEE square roots per particle: +0.7 % on the double
// gyre at 8192^2 with identical config-1 parity figures (profiles/r3_ab_tile_hinit.txt).
#ifndef B200CS_HINIT_FAST
#define B200CS_HINIT_FAST (B200CS_CTRL_FAST && B200CS_LEAN2)
#endif
#ifndef B200CS_SYNC_EVERY
#define B200CS_SYNC_EVERY 8
#endif

namespace b200cs {

struct StepCounts {
    int accepted = 0;
    int rejected = 0;  // rejected attempts (including those before the first accepted step)
    int dense = 0;     // accepted steps that needed the three extra dense-output stages
};

// The controller's non-trivial literals in the constant bank (B200CS_CTL_CONSTS): a double that is not a
// short immediate costs two UMOV per use on sm_100 (one per 32-bit half) and two uniform registers
// that ptxas then lacks for the tableau: with these seven literals as LDCU.64 loads the
// double-gyre attempt went from 293 to 243 non-FP64 instructions and from 62 bytes of spills to 4
// (the uniform-register file had been overflowing into local memory and R2UR moves):
// 1170.6 -> 1192.3 M points/s at 8192^2, bit-identical results (profiles/r3_ab_nonfp64.txt).
struct __align__(16) CtlConsts {
    double tenth, uround, one01, hundredth, safe, facmin, tiny, pad;
};
static __constant__ CtlConsts kCtl = {0.1, 2.3e-16, 1.01, 0.01, 0.9, 0.333, 1.0e-280, 0.0};
#ifndef B200CS_CTL_CONSTS
#define B200CS_CTL_CONSTS 1
#endif
#if B200CS_CTL_CONSTS
#define B2_CTL(field, literal) (::b200cs::kCtl.field)
#else
#define B2_CTL(field, literal) (literal)
#endif

// Optional part of the RHS interface (detected, so the other flows need not mention it):
//   Rhs::kAuxAffine     time_part_affine<M>(x, h, c[M], aux) evaluates the time-only part at the
//                       times x + c_m h without forming them (the phase is affine in c_m).
// (Tried and measured equal: leaving the constant amplitudes of the double-gyre RHS to the
// integrator as hs_i = h * scale_i -- two multiplications fewer per stage, 993.6 vs 998.1 M
// points/s at 8192^2, profiles/r1d_ab_variants.txt -- so the slopes stay in true units.)
template <class T, class = void>
struct rhs_affine : std::false_type {};
template <class T>
struct rhs_affine<T, std::void_t<decltype(T::kAuxAffine)>> : std::bool_constant<T::kAuxAffine> {};

// Where the stage slopes K[1..16][N] live.  RegSlopes: registers (the analytic flows: the attempt loop
// is FP64-bound and every slope is a compile-time-indexed register).  SmemSlopes: shared memory,
// element (s, i) of thread t at base[(s N + i) blockDim + t] (conflict-free); for the flows whose
// right-hand side is large (64-tap spline) the ~50 registers the slopes would pin are worth more
// as load / FMA registers of the RHS, and the ~170 shared-memory accesses per attempt are noise
// next to its 768 coefficient loads.
template <int N>
struct RegSlopes {
    static constexpr bool kDirect = true;
    double a[17][N];
    __device__ __forceinline__ double (&operator[](int s))[N] { return a[s]; }
    __device__ __forceinline__ const double (&operator[](int s) const)[N] { return a[s]; }
};
template <int N>
struct SmemSlopes {
    static constexpr bool kDirect = false;
    double *base;   // shared memory, already offset by the thread index
    int stride;     // threads per block
    // reads are volatile: without that ptxas forwards every stored slope in registers (the addresses are
    // provably thread-private) and the point of parking them -- fewer live registers -- is lost
    struct Ref {
        double *q;
        __device__ __forceinline__ operator double() const { return *reinterpret_cast<volatile double *>(q); }
        __device__ __forceinline__ Ref &operator=(double v) {
            *q = v;
            return *this;
        }
        __device__ __forceinline__ Ref &operator=(const Ref &o) { return *this = (double)o; }
    };
    struct Row {
        double *p;
        int stride;
        __device__ __forceinline__ Ref operator[](int i) const { return Ref{p + i * stride}; }
    };
    __device__ __forceinline__ Row operator[](int s) const { return Row{base + s * N * stride, stride}; }
};

namespace detail {

// K[S] = rhs(aux, t, yy): straight into the register row, or through a temporary into shared memory
template <class Rhs, class KS, int N>
__device__ __forceinline__ void eval_to(const Rhs &rhs, double aux, double t, const double (&yy)[N], KS &K, int S) {
    if constexpr (KS::kDirect) {
        rhs.eval(aux, t, yy, K[S]);
    } else {
        double tmp[N];
        rhs.eval(aux, t, yy, tmp);
#pragma unroll
        for (int i = 0; i < N; ++i) K[S][i] = tmp[i];
    }
}

// a*b + c: one FMA in the product build, two roundings (the reference's arithmetic) in the strict one
__device__ __forceinline__ double mad(double a, double b, double c) {
#if B200CS_STRICT_INT
    return __dadd_rn(__dmul_rn(a, b), c);
#else
    return fma(a, b, c);
#endif
}

// err^(1/8) with three correctly rounded square roots (libm pow in the reference; the difference
// is <= 1 ulp and only scales the next step size)
#if B200CS_LEAN2
// three square roots as x * rsqrt(x) (MUFU seed + one Newton step each, <= 2 ulp) instead of three
// correctly rounded ones.  NaN must stay NaN (a NaN error norm has to reach the controller's
// fmin / fmax as NaN so that the step is rejected with h/3, ending in B200CS_ST_HSMALL after a
// few dozen attempts, exactly as with sqrt); +inf gives NaN, which the same fmin treats like +inf.
__device__ __forceinline__ double sqrt_fast(double x) { return x == 0.0 ? 0.0 : x * rsqrt(x); }
__device__ __forceinline__ double pow_eighth(double x) { return sqrt_fast(sqrt_fast(sqrt_fast(x))); }
#elif B200CS_STRICT_INT
__device__ __forceinline__ double pow_eighth(double x) { return pow(x, 0.125); }  // libm pow, as the reference
#else
__device__ __forceinline__ double pow_eighth(double x) { return sqrt(sqrt(sqrt(x))); }
#endif
#if B200CS_CTRL_FAST
// 1/x for a normal positive x: MUFU.RCP64H seed (~2^-22) and two Newton steps
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}
// x^(-1/8) for x > 0: z0 from three chained MUFU.RSQ64H seeds (x^-1/2 -> x^-1/4 -> x^-1/8, ~1e-6),
// then Newton on z^-8 = x:  z <- z + z (1 - x z^8) / 8   (quadratic, two steps -> rounding level)
__device__ __forceinline__ double inv_eighth_root(double x) {
    double s1, s2, z;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s1) : "d"(x));
    const double t1 = x * s1;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s2) : "d"(t1));
    const double t2 = t1 * s2;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(t2));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
        z = fma(z * fma(-x, z8, 1.0), 0.125, z);
    }
    return z;
}
#endif
// The one-Newton-step forms (B200CS_CTRL_ONE_NEWTON): 1/sk to ~6e-14 and err^(-1/8) to ~5e-12 relative.
// The scaled error of a step moves by 6e-14 relative (an accept / reject test err <= 1 flips for
// ~1e-4 particles of a 2.7e8-particle grid), the next step size by 5e-12 relative, which moves a
// step's result by 8 x (local error ~1e-6) x 5e-12 = 4e-17: far below the 1-ulp noise of the RHS.
__device__ __forceinline__ double rcp_fast1(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double inv_eighth_root1(double x) {
    double s1, s2, z;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s1) : "d"(x));
    const double t1 = x * s1;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s2) : "d"(t1));
    const double t2 = t1 * s2;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(z) : "d"(t2));
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    return fma(z * fma(-x, z8, 1.0), 0.125, z);
}
// a / c for a compile-time constant c: q = a*rc, one FMA residual correction -> the correctly
// rounded quotient in 3 FP64 instructions (see tensor_kernels.cu, div_const)
__device__ __forceinline__ double div_const(double a, double c, double rc) {
    const double q = a * rc;
    return fma(fma(-q, c, a), rc, q);
}

// yy = y + h * sum_j a(S,j) K_j   (j ascending, zero entries skipped at compile time)
template <int S, int N, class KS, int... J>
__device__ __forceinline__ void stage_arg(double (&yy)[N], const double (&y)[N], double h,
                                          const KS &K, std::integer_sequence<int, J...>) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double acc = 0.0;
        ((dop::a(S, J + 1) != 0.0 ? (void)(acc = mad(dop::kTab.a[S][J + 1], K[J + 1][i], acc)) : (void)0), ...);
        yy[i] = mad(h, acc, y[i]);
    }
}

// stage S: K_S = f(x + c_S h, y + h sum a_Sj K_j); `aux` is the precomputed time-only part
template <int S, class Rhs, int N, class KS>
__device__ __forceinline__ void do_stage(const Rhs &rhs, double aux, double x, double h, const double (&y)[N],
                                         KS &K) {
    double yy[N];
    stage_arg<S, N>(yy, y, h, K, std::make_integer_sequence<int, S - 1>{});
    eval_to(rhs, aux, mad(dop::kTab.c[S], h, x), yy, K, S);
}

// time-only part of the RHS at the stage times S0 .. S0+M-1 of the step (x, h), M chains at once
template <int S0, int M, class Rhs>
__device__ __forceinline__ void stage_aux(const Rhs &rhs, double x, double h, double (&aux)[17]) {
    if constexpr (Rhs::kAux != 0) {
        double t[M], a[M];
        if constexpr (rhs_affine<Rhs>::value) {
#pragma unroll
            for (int m = 0; m < M; ++m) t[m] = dop::kTab.c[S0 + m];
            rhs.template time_part_affine<M>(x, h, t, a);
        } else {
#pragma unroll
            for (int m = 0; m < M; ++m) t[m] = mad(dop::kTab.c[S0 + m], h, x);
            rhs.template time_part<M>(t, a);
        }
#pragma unroll
        for (int m = 0; m < M; ++m) aux[S0 + m] = a[m];
    }
}

template <int R, int N, class KS, int... J>
__device__ __forceinline__ double dense_row(const KS &K, int i, std::integer_sequence<int, J...>) {
    double acc = 0.0;
    ((dop::d(R, J + 1) != 0.0 ? (void)(acc = mad(dop::kTab.d[R][J + 1], K[J + 1][i], acc)) : (void)0), ...);
    return acc;
}

}  // namespace detail

// Output sink for the n-time variant: called once per emitted row k (0 < k < n-1 via dense
// output, k == n-1 with the end point; row 0 is written by the caller).
template <int N>
struct NoSink {
    __device__ __forceinline__ void operator()(int, const double (&)[N]) const {}
};

// First slope K[1] = f(x, y) and the initial step size of Hairer's hinit (iord = 8); returns h.
template <class Rhs, int N, class KS>
__device__ __forceinline__ double dop853_start(const Rhs &rhs, double x, const double (&y)[N], double rtol,
                                               double atol, double hmax, double posneg, KS &K) {
    double h;
    {
        double t1[1] = {x}, a1[1] = {0.0};
        if constexpr (Rhs::kAux != 0) rhs.template time_part<1>(t1, a1);
        detail::eval_to(rhs, a1[0], x, y, K, 1);
    }
    // ---- hinit (iord = 8)
#if B200CS_HINIT_FAST
    {   // the same formulas with the controller's reciprocal / rsqrt kernels instead of nine IEEE
        // divisions and three IEEE square roots (once per particle, ~1 % of its FP64 instructions);
        // NaN inputs end in h = +-hmax exactly as with the IEEE operations (fmin drops the NaN)
        double dnf = 0.0, dny = 0.0, rsk[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            rsk[i] = detail::rcp_fast(detail::mad(rtol, fabs(y[i]), atol));
            const double a = K[1][i] * rsk[i], b = y[i] * rsk[i];
            dnf = detail::mad(a, a, dnf);
            dny = detail::mad(b, b, dny);
        }
        h = (dnf <= 1.0e-10 || dny <= 1.0e-10) ? 1.0e-6 : detail::sqrt_fast(dny * detail::rcp_fast(dnf)) * 0.01;
        h = fmin(h, hmax) * posneg;
        double y1[N], f1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = detail::mad(h, K[1][i], y[i]);
        double t1[1] = {x + h}, a1[1] = {0.0};
        if constexpr (Rhs::kAux != 0) rhs.template time_part<1>(t1, a1);
        rhs.eval(a1[0], x + h, y1, f1);
        double der2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double a = (f1[i] - K[1][i]) * rsk[i];
            der2 = detail::mad(a, a, der2);
        }
        der2 = detail::sqrt_fast(der2) * detail::rcp_fast(fabs(h));   // |sqrt(der2) / h|
        const double der12 = fmax(der2, detail::sqrt_fast(dnf));
        const double h1 = (der12 <= 1.0e-15) ? fmax(1.0e-6, fabs(h) * 1.0e-3)
                                             : detail::rcp_fast(detail::inv_eighth_root(0.01 * detail::rcp_fast(der12)));
        h = fmin(100.0 * fabs(h), fmin(h1, hmax)) * posneg;
    }
#else
    {
        double dnf = 0.0, dny = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = detail::mad(rtol, fabs(y[i]), atol);
            const double a = K[1][i] / sk, b = y[i] / sk;
            dnf = detail::mad(a, a, dnf);
            dny = detail::mad(b, b, dny);
        }
        h = (dnf <= 1.0e-10 || dny <= 1.0e-10) ? 1.0e-6 : sqrt(dny / dnf) * 0.01;
        h = fmin(h, hmax) * posneg;
        double y1[N], f1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = detail::mad(h, K[1][i], y[i]);
        double t1[1] = {x + h}, a1[1] = {0.0};
        if constexpr (Rhs::kAux != 0) rhs.template time_part<1>(t1, a1);
        rhs.eval(a1[0], x + h, y1, f1);
        double der2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sk = detail::mad(rtol, fabs(y[i]), atol);
            const double a = (f1[i] - K[1][i]) / sk;
            der2 = detail::mad(a, a, der2);
        }
        der2 = sqrt(der2) / h;
        const double der12 = fmax(fabs(der2), sqrt(dnf));
        const double h1 = (der12 <= 1.0e-15) ? fmax(1.0e-6, fabs(h) * 1.0e-3)
                                             : detail::pow_eighth(0.01 / der12);
        h = fmin(100.0 * fabs(h), fmin(h1, hmax)) * posneg;
    }
#endif
    return h;
}

// Work feeder of the queue kernels (flowmap_kernel.cuh): lanes that have finished their particle
// hand in the result and take the next pre-initialised one at the top of the attempt loop, so a
// warp is not held by its slowest lane.  NoFeeder = one particle per lane, as everywhere else.
struct NoFeeder {
    static constexpr bool kActive = false;
};

// Integrates dy/ds = rhs(s, y) from s = x0 to s = xend.
//   n_out  == 0 : only the final state is wanted (DENSE must be false)
//   n_out  >= 2 : output times are t_k = p0 * (t0 + k*step), k = 0..n_out-1 (last = p0*(t0+T)),
//                 rows 1..n_out-1 go to `sink`
// Returns B200CS_ST_OK / _NMAX / _HSMALL; y holds the state reached.
template <bool DENSE, bool LOCKSTEP, class Rhs, int N, class Sink, class KS, class Feeder = NoFeeder>
__device__ __forceinline__ int dop853_integrate(const Rhs &rhs, bool active, double (&y)[N], double x0, double xend,
                                                double rtol, double atol, int n_out, double out_p0,
                                                double out_t0, double out_step, Sink &&sink,
                                                StepCounts &cnt, KS K, Feeder &&feeder = Feeder{}) {
    constexpr bool kFed = std::remove_reference_t<Feeder>::kActive;
    static_assert(!(kFed && DENSE), "the work feeder serves the final-time kernels");
    constexpr double kSafe = 0.9, kFacc1 = 1.0 / 0.333, kFacc2 = 1.0 / 6.0, kURound = 2.3e-16;
    constexpr int kNmax = 100000;
    constexpr int kSyncEvery = B200CS_SYNC_EVERY;
    // K[1..12] stage slopes, K[13] FSAL slope, K[14..16] dense-output stages (registers or shared memory)
    double aux[17];   // time-only part of the RHS per stage (flows that have one)
    double x = x0;
    const double posneg = (xend - x0) < 0.0 ? -1.0 : 1.0;
    const double hmax = fabs(xend - x0);
    bool last = false, reject = false;
    int nstep = 0;
    int iout = 1;  // next output row
    // output time of row k, bit-identical to params[0]*np.linspace(t0, t0+T, n)[k]
    auto t_out = [&](int k) { return __dmul_rn(out_p0, __dadd_rn(out_t0, __dmul_rn((double)k, out_step))); };
    double tnext = 0.0;
    if (DENSE) tnext = t_out(1);
    // LOCKSTEP: every thread of the block calls this function (inactive ones idle) and the block
    // meets at a barrier before each step attempt, so all its warps walk the large unrolled body
    // together and share its instruction-cache footprint.
    bool alive = active && !kFed;
    double h = 0.0;
    int status = active ? B200CS_ST_OK : B200CS_ST_MASKED;

    if (alive && !kFed) h = dop853_start(rhs, x, y, rtol, atol, hmax, posneg, K);

    for (int it = 0;; ++it) {
        if constexpr (kFed && LOCKSTEP) {
            // queue kernel in lockstep: warps refill on their own, the block re-aligns every kSyncEvery
            // attempts and leaves together once no lane is alive and every warp has seen the queue empty
            if (__any_sync(0xffffffffu, !alive)) {
                if (feeder.refill(alive, y, h, K[1], status, cnt)) {
                    x = x0;
                    last = false;
                    reject = false;
                    nstep = 0;
                    status = B200CS_ST_OK;
                    cnt = StepCounts{};
                    alive = true;
                }
            }
            if ((it % B200CS_QSYNC_EVERY) == 0 && !__syncthreads_or((alive || !feeder.drained()) ? 1 : 0)) break;
        } else if (LOCKSTEP) {
            // re-align the block every kSyncEvery attempts (warps drift apart only slowly); the
            // loop is left at a barrier, by all threads together, once nobody is alive
            if ((it % kSyncEvery) == 0 && !__syncthreads_or(alive ? 1 : 0)) break;
        } else if constexpr (kFed && !LOCKSTEP) {
            if (__any_sync(0xffffffffu, !alive)) {
                // all lanes call (the fetch is warp-cooperative); a lane that takes a particle comes back
                // alive with y, h and the first slope of a fresh integration
                if (feeder.refill(alive, y, h, K[1], status, cnt)) {
                    x = x0;
                    last = false;
                    reject = false;
                    nstep = 0;
                    status = B200CS_ST_OK;
                    cnt = StepCounts{};
                    alive = true;
                }
                if (!__any_sync(0xffffffffu, alive)) {
                    if (feeder.drained()) break;
                    continue;
                }
            }
        } else if (!alive) {
            break;
        }
        if (alive) do {
        if (nstep > kNmax) { status = B200CS_ST_NMAX; alive = false; break; }
        if (B2_CTL(tenth, 0.1) * fabs(h) <= fabs(x) * B2_CTL(uround, kURound)) { status = B200CS_ST_HSMALL; alive = false; break; }
        if ((x + B2_CTL(one01, 1.01) * h - xend) * posneg > 0.0) {
            h = xend - x;
            last = true;
        }
        ++nstep;
        const double xph = x + h;
        // the stage times x + c_s h are known now: the time-only part of the RHS is evaluated for
        // four stages at a time (four independent chains), off the stages' critical path
        detail::stage_aux<2, 4>(rhs, x, h, aux);
        detail::do_stage<2>(rhs, aux[2], x, h, y, K);
        detail::do_stage<3>(rhs, aux[3], x, h, y, K);
        detail::do_stage<4>(rhs, aux[4], x, h, y, K);
        detail::do_stage<5>(rhs, aux[5], x, h, y, K);
        detail::stage_aux<6, 4>(rhs, x, h, aux);
        detail::do_stage<6>(rhs, aux[6], x, h, y, K);
        detail::do_stage<7>(rhs, aux[7], x, h, y, K);
        detail::do_stage<8>(rhs, aux[8], x, h, y, K);
        detail::do_stage<9>(rhs, aux[9], x, h, y, K);
        detail::stage_aux<10, 3>(rhs, x, h, aux);  // c12 = 1: aux[12] also serves the FSAL slope
        detail::do_stage<10>(rhs, aux[10], x, h, y, K);
        detail::do_stage<11>(rhs, aux[11], x, h, y, K);
        {   // stage 12 is evaluated at x + h exactly
            double yy[N];
            detail::stage_arg<12, N>(yy, y, h, K, std::make_integer_sequence<int, 11>{});
            detail::eval_to(rhs, aux[12], xph, yy, K, 12);
        }
        // 8th-order slope, candidate state, and the two embedded error estimates
        double y5[N];
        double err = 0.0, err2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = dop::kTab.b[1] * K[1][i];
            s = detail::mad(dop::kTab.b[6], K[6][i], s);
            s = detail::mad(dop::kTab.b[7], K[7][i], s);
            s = detail::mad(dop::kTab.b[8], K[8][i], s);
            s = detail::mad(dop::kTab.b[9], K[9][i], s);
            s = detail::mad(dop::kTab.b[10], K[10][i], s);
            s = detail::mad(dop::kTab.b[11], K[11][i], s);
            s = detail::mad(dop::kTab.b[12], K[12][i], s);
            y5[i] = detail::mad(h, s, y[i]);
#if B200CS_LEAN_TAIL >= 2
            // max(|y|, |y5|) as one compare and a select (fmax's NaN fix-up is an extra instruction; a NaN
            // y5 is ignored here exactly as fmax ignores it, and err is NaN through e5 then anyway)
            const double ay = fabs(y[i]), ay5 = fabs(y5[i]);
            const double sk = detail::mad(rtol, ay5 > ay ? ay5 : ay, atol);
#else
            const double sk = detail::mad(rtol, fmax(fabs(y[i]), fabs(y5[i])), atol);
#endif
            double e3 = detail::mad(-dop::kTab.bhh[0], K[1][i], s);
            e3 = detail::mad(-dop::kTab.bhh[1], K[9][i], e3);
            e3 = detail::mad(-dop::kTab.bhh[2], K[12][i], e3);
#if B200CS_CTRL_FAST
#if B200CS_CTRL_ONE_NEWTON
            const double rsk = detail::rcp_fast1(sk);
#else
            const double rsk = detail::rcp_fast(sk);  // sk = atol + rtol |y| is a normal positive number
#endif
            e3 *= rsk;
#elif B200CS_LEAN
            const double rsk = 1.0 / sk;  // one reciprocal serves both estimates (<= 1 ulp from e/sk)
            e3 *= rsk;
#else
            e3 /= sk;
#endif
            err2 = detail::mad(e3, e3, err2);
            double e5 = dop::kTab.er[1] * K[1][i];
            e5 = detail::mad(dop::kTab.er[6], K[6][i], e5);
            e5 = detail::mad(dop::kTab.er[7], K[7][i], e5);
            e5 = detail::mad(dop::kTab.er[8], K[8][i], e5);
            e5 = detail::mad(dop::kTab.er[9], K[9][i], e5);
            e5 = detail::mad(dop::kTab.er[10], K[10][i], e5);
            e5 = detail::mad(dop::kTab.er[11], K[11][i], e5);
            e5 = detail::mad(dop::kTab.er[12], K[12][i], e5);
#if B200CS_LEAN || B200CS_CTRL_FAST
            e5 *= rsk;
#else
            e5 /= sk;
#endif
            err = detail::mad(e5, e5, err);
        }
        double deno = detail::mad(B2_CTL(hundredth, 0.01), err2, err);
        if (deno <= 0.0) deno = 1.0;
#if B200CS_LEAN
        err = fabs(h) * err * rsqrt(deno * (double)N);
#else
        err = fabs(h) * err * sqrt(1.0 / (deno * (double)N));
#endif
#if B200CS_CTRL_FAST
        // g = 0.9 err^(-1/8) = 1 / (fac11 / safe); err == 0 (or below the seed's range) -> +inf, i.e. the
        // growth clamp; a NaN err stays NaN and fmax / fmin then pick the 1/3 of a rejected step
#if B200CS_CTRL_ONE_NEWTON
        double g = B2_CTL(safe, kSafe) * detail::inv_eighth_root1(err);
#else
        double g = kSafe * detail::inv_eighth_root(err);
#endif
        if (err <= B2_CTL(tiny, 1.0e-280)) g = __longlong_as_double(0x7ff0000000000000LL);
        double hnew = h * fmin(1.0 / kFacc2, fmax(B2_CTL(facmin, 1.0 / kFacc1), g));
#else
        const double fac11 = detail::pow_eighth(err);
#if B200CS_LEAN2
        const double fac11s = detail::div_const(fac11, kSafe, 1.0 / kSafe);
#else
        const double fac11s = fac11 / kSafe;
#endif
        const double fac = fmax(kFacc2, fmin(kFacc1, fac11s));  // beta = 0
        double hnew = h / fac;
#endif
#if B200CS_FSAL_ALWAYS
        detail::eval_to(rhs, aux[12], xph, y5, K, 13);  // A/B: FSAL slope evaluated before the accept test
#endif
        if (err <= 1.0) {
            // ---- accepted
            ++cnt.accepted;
#if !B200CS_FSAL_ALWAYS
#if B200CS_LEAN_TAIL
            // final-time kernels: the first-same-as-last slope goes straight into K[1] (nothing reads the
            // old one any more; only the dense-output coefficients need both)
            detail::eval_to(rhs, aux[12], xph, y5, K, DENSE ? 13 : 1);
#else
            detail::eval_to(rhs, aux[12], xph, y5, K, 13);  // first-same-as-last slope, same time as stage 12
#endif
#endif
            if (DENSE) {
                if (iout < n_out - 1 && (tnext - xph) * posneg <= 0.0) {
                    ++cnt.dense;
                    double rc[8][N];
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        rc[0][i] = y[i];
                        const double ydiff = y5[i] - y[i];
                        rc[1][i] = ydiff;
                        const double bspl = detail::mad(h, K[1][i], -ydiff);
                        rc[2][i] = bspl;
                        rc[3][i] = ydiff - h * K[13][i] - bspl;
                    }
                    detail::stage_aux<14, 3>(rhs, x, h, aux);
                    detail::do_stage<14>(rhs, aux[14], x, h, y, K);
                    detail::do_stage<15>(rhs, aux[15], x, h, y, K);
                    detail::do_stage<16>(rhs, aux[16], x, h, y, K);
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        rc[4][i] = h * detail::dense_row<4, N>(K, i, std::make_integer_sequence<int, 16>{});
                        rc[5][i] = h * detail::dense_row<5, N>(K, i, std::make_integer_sequence<int, 16>{});
                        rc[6][i] = h * detail::dense_row<6, N>(K, i, std::make_integer_sequence<int, 16>{});
                        rc[7][i] = h * detail::dense_row<7, N>(K, i, std::make_integer_sequence<int, 16>{});
                    }
                    // contd8 at the output time tq (Hairer's nested form)
#if B200CS_DENSE_RCP
                    const double rh = 1.0 / h;   // one division per step instead of one per output row
#endif
                    auto dense_at = [&](double tq, double (&yo)[N]) {
#if B200CS_DENSE_RCP
                        const double th = (tq - x) * rh, th1 = 1.0 - th;
#else
                        const double th = (tq - x) / h, th1 = 1.0 - th;
#endif
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            double v = th * rc[7][i];
                            v = th1 * (rc[6][i] + v);
                            v = th * (rc[5][i] + v);
                            v = th1 * (rc[4][i] + v);
                            v = th * (rc[3][i] + v);
                            v = th1 * (rc[2][i] + v);
                            v = th * (rc[1][i] + v);
                            yo[i] = rc[0][i] + v;
                        }
                    };
                    do {
                        double yo[N];
                        dense_at(tnext, yo);
                        sink(iout, yo);
                        ++iout;
                        tnext = t_out(iout);
                    } while (iout < n_out - 1 && (tnext - xph) * posneg <= 0.0);
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
#if B200CS_LEAN_TAIL
                if (DENSE) K[1][i] = K[13][i];
#else
                K[1][i] = K[13][i];
#endif
                y[i] = y5[i];
            }
            x = xph;
#if B200CS_LEAN_TAIL
            alive = !last;   // one way out of the attempt: the step size below is simply not used any more
#else
            if (last) { alive = false; break; }
#endif
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
            reject = false;
        } else {
            // ---- rejected
#if B200CS_CTRL_FAST
            hnew = h * fmax(B2_CTL(facmin, 1.0 / kFacc1), g);
#else
            hnew = h / fmin(kFacc1, fac11s);
#endif
            reject = true;
            last = false;
            ++cnt.rejected;
        }
        h = hnew;
        } while (0);
    }
    if (DENSE && status == B200CS_ST_OK && n_out >= 2) sink(n_out - 1, y);
    if (DENSE && n_out >= 2 && (status == B200CS_ST_NMAX || status == B200CS_ST_HSMALL)) {
        // the integration gave up: the rows it never reached are NaN, not whatever the output
        // buffer held (the reference's numbalsoda leaves them unspecified and ignores `success`)
        double bad[N];
#pragma unroll
        for (int i = 0; i < N; ++i) bad[i] = __longlong_as_double(0x7ff8000000000000LL);
        for (int k = iout; k < n_out; ++k) sink(k, bad);
    }
    return status;
}

}  // namespace b200cs
