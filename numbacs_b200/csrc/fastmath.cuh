// fastmath.cuh -- lean FP64 sin / cos / expm1 for the right-hand sides.
//
// The RHS of the analytic flows is dominated by transcendentals (flows.py:1152-1158: five sin/cos
// per double-gyre call; flows.py:1189-1213: cosh, tanh and six sin/cos per Bickley call).  CUDA's
// libm versions are accurate but wrap ~24 FP64 instructions in ~60 integer / uniform / branch
// instructions (immediates materialised with UMOV pairs, inf/NaN checks, a Payne-Hanek slow path
// behind a CALL).  On B200 the FP64 pipe issues one warp instruction every two cycles per SM
// sub-partition and an FP64 instruction cannot take a constant-bank operand directly, so that
// overhead -- not FP64 throughput -- would bound the kernels.  Two families live here:
//   * the "wide" kernels (sinpi12_v, sin_wide_v, sincos_wide, expm1_neg; round 1c-1f): reduce to
//     half a period only and evaluate ONE odd / even polynomial there, so a sine is 12-14 FP64
//     instructions plus a sign flip and every coefficient is a warp-uniform constant-bank load.
//     These are what the double-gyre, Bickley-jet and ABC right-hand sides use;
//   * the parity-selected kernels (sin_v, sinpi_v, sincos_fast; rounds 1a/1b): reduce to a
//     quarter period with the classical minimax kernels on [-pi/4, pi/4] (< 1 ulp) and pick the
//     sine or cosine polynomial per lane through an indexed constant load.  One FP64 instruction
//     fewer per sine but ~15 more issue slots; still used by cos_fast (one call per spline RHS)
//     and kept for A/B builds (B200CS_SINPI_WIDE=0 etc., tools/build_variant.py).
// All of them evaluate M independent arguments side by side (M dependency chains, each constant
// fetched once) and hand arguments outside their exact-reduction range (|x| >= 1e5, for sin(pi u)
// |u| >= 2^50, NaN, inf) to libm, out of line.  Accuracy is measured on the CPU from the constants
// in this file (tests/test_trig_poly_cpu.py, tools/fit_trig_poly.py).
#pragma once
#include <cuda_runtime.h>

namespace b200cs {

struct __align__(16) TrigConsts {
    double two_over_pi, magic, p1, p2, p3, pad;
    double s[6];      // sin kernel: r + r^3 (s0 + z s1 + ... + z^5 s5), z = r^2
    double c[6];      // cos kernel: 1 - z/2 + z^2 (c0 + z c1 + ... + z^5 c5)
    double sc[6][2];  // the same coefficients interleaved: sc[k][0] = s[k], sc[k][1] = c[k]
};

static __constant__ TrigConsts kTrig = {
    0.63661977236758138,      // 2/pi
    6755399441055744.0,       // 1.5 * 2^52: adding it rounds to the nearest integer
    1.5707963267948966,       // pi/2 = p1 + p2 + p3 (+ O(1e-48)); p1 = fl(pi/2)
    6.123233995736757e-17,    // 0x3c91a62633145c00: trailing zero bits keep k*p2 exact for |k| < 2^16
    8.478427660368898e-32,    // 0x397b839a252049c0
    0.0,
    {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
     2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10},
    {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
     -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11},
    {{-1.66666666666666324348e-01, 4.16666666666666019037e-02},
     {8.33333333332248946124e-03, -1.38888888888741095749e-03},
     {-1.98412698298579493134e-04, 2.48015872894767294178e-05},
     {2.75573137070700676789e-06, -2.75573143513906633035e-07},
     {-2.50507602534068634195e-08, 2.08757232129817482790e-09},
     {1.58969099521155010221e-10, -1.13596475577881948265e-11}}};

// libm fall-back for huge / non-finite arguments, kept out of line so that its Payne-Hanek code
// is not replicated at every call site
static __device__ __noinline__ double2 sincos_slow(double x) {
    double2 r;
    sincos(x, &r.x, &r.y);
    return r;
}

// Round 2, last session: the FP64 pipe of the integration kernels idles ~15 % of the time although
// the FP64 instruction stream could fill it, and what moves the needle is the NON-FP64 issue slots
// between them (profiles/r3_ab_nonfp64.txt, double gyre at 8192^2):
//   B200CS_FLIP_ADD    the sign flip (-1)^k as ONE integer multiply-add on the high word
//                      (k * 2^31 + hi: the carry leaves the word) instead of a shift and a LOP3:
//                      320 -> 295 non-FP64 instructions per attempt, 1150.6 -> 1170.8 M points/s,
//                      bit-identical;
//   B200CS_FLIP_EARLY  the flip applied to r (the polynomial is odd) instead of to the result, so it
//                      leaves the Horner chain's critical path.  Value 2 (default) keeps the final
//                      product r * Q rounded on its own (__dmul_rn), exactly as when the integer flip
//                      fenced it: bit-identical results.  Value 1 lets it contract with the sum that
//                      follows (794 -> 782 FP64 instructions, no faster) -- and breaks the wall identity
//                      S+ + S- == 0 of the double gyre (the unrounded product leaves a 1e-17 residue
//                      where the rounded ones cancel), which moves wall particles' first step and
//                      with it the float32 goldens: rejected.
#ifndef B200CS_FLIP_ADD
#define B200CS_FLIP_ADD 1
#endif
#ifndef B200CS_FLIP_EARLY
#define B200CS_FLIP_EARLY 2
#endif
// flips the sign of d when bit is 1
__device__ __forceinline__ double flip_sign(double d, int bit) {
#if B200CS_FLIP_ADD
    // adding 2^31 to the high word flips bit 31 (the carry leaves the word): one IMAD (bit * 2^31 + hi)
    // instead of a shift and a LOP3
    return __hiloint2double((int)((unsigned)__double2hiint(d) + ((unsigned)bit << 31)), __double2loint(d));
#else
    return __hiloint2double(__double2hiint(d) ^ (bit << 31), __double2loint(d));
#endif
}

// r = x - k pi/2 with k = rint(x 2/pi) (|x| < 1e5); q = k mod 2^32
__device__ __forceinline__ double trig_reduce(double x, int &q) {
    const double t = fma(x, kTrig.two_over_pi, kTrig.magic);
    q = __double2loint(t);
    const double k = t - kTrig.magic;
    double r = fma(-k, kTrig.p1, x);
#ifdef B200CS_CW3
    r = fma(-k, kTrig.p2, r);
    return fma(-k, kTrig.p3, r);
#else
    // two Cody-Waite terms: p1 + p2 carries pi/2 to 107 bits, so the dropped k*p3 (< 1e-26 for
    // |k| < 1e5) only matters when r itself is below ~1e-10, i.e. within rounding of a zero of
    // sin/cos, where it changes the result by a few ulps of an O(1e-16) number
    return fma(-k, kTrig.p2, r);
#endif
}

// s[m] = sin(x[m] + shift * pi/2) for M independent arguments (shift = 1 gives cos).  ONE
// polynomial per argument: its coefficients are picked by the quadrant parity through an indexed
// constant-bank load (at most two distinct addresses per warp), so there is no branch, no
// divergence and no second code copy.
// |x| < 1e5 as an integer compare on the high word (1e5 = 0x40F86A00'00000000; false for NaN / inf):
// keeps the range check off the FP64 pipe, which is the kernel's bottleneck
__device__ __forceinline__ bool trig_in_range(double x) {
    return (unsigned)(__double2hiint(x) & 0x7fffffff) < 0x40F86A00u;
}

template <int M, int SHIFT = 0>
__device__ __forceinline__ void sin_v(const double (&x)[M], double (&s)[M]) {
    bool slow = false;
#pragma unroll
    for (int m = 0; m < M; ++m) slow |= !trig_in_range(x[m]);  // also catches NaN / inf
    if (slow) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const double2 sc = sincos_slow(x[m]);
            s[m] = SHIFT ? sc.y : sc.x;
        }
        return;
    }
    int q[M];
    double r[M], z[M], p[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        r[m] = trig_reduce(x[m], q[m]);
        q[m] += SHIFT;
        z[m] = r[m] * r[m];
        p[m] = kTrig.sc[5][q[m] & 1];
    }
#pragma unroll
    for (int k = 4; k >= 0; --k)
#pragma unroll
        for (int m = 0; m < M; ++m) p[m] = fma(p[m], z[m], kTrig.sc[k][q[m] & 1]);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const bool par = q[m] & 1;
        const double mm = z[m] * (par ? z[m] : r[m]);             // z^2 (cos kernel) or z r (sin kernel)
        const double base = par ? fma(z[m], -0.5, 1.0) : r[m];    // 1 - z/2 or r
        s[m] = flip_sign(fma(mm, p[m], base), (q[m] >> 1) & 1);
    }
}

// s[m] = sin(pi * u[m]).  The reduction is exact: k = rint(2u), r = u - k/2 with |r| <= 1/4 is
// representable (a difference of nearby doubles), so the only rounding before the polynomial is
// the single product pi*r -- three FP64 instructions fewer than forming pi*u and reducing it by
// pi/2 in three Cody-Waite steps, and a smaller absolute error (no rounding of pi*u at |pi*u| ~ 6).
template <int M>
__device__ __forceinline__ void sinpi_v(const double (&u)[M], double (&s)[M]) {
    bool slow = false;
#pragma unroll
    for (int m = 0; m < M; ++m) slow |= !trig_in_range(u[m]);
    if (slow) {
#pragma unroll
        for (int m = 0; m < M; ++m) s[m] = sincos_slow(3.141592653589793 * u[m]).x;
        return;
    }
    int q[M];
    double r[M], z[M], p[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const double t = fma(u[m], 2.0, kTrig.magic);
        q[m] = __double2loint(t);
        const double k = t - kTrig.magic;
        r[m] = fma(k, -0.5, u[m]) * 3.141592653589793;
        z[m] = r[m] * r[m];
        p[m] = kTrig.sc[5][q[m] & 1];
    }
#pragma unroll
    for (int k = 4; k >= 0; --k)
#pragma unroll
        for (int m = 0; m < M; ++m) p[m] = fma(p[m], z[m], kTrig.sc[k][q[m] & 1]);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const bool par = q[m] & 1;
        const double mm = z[m] * (par ? z[m] : r[m]);
        const double base = par ? fma(z[m], -0.5, 1.0) : r[m];
        s[m] = flip_sign(fma(mm, p[m], base), (q[m] >> 1) & 1);
    }
}

// ---- "wide" kernels: ONE odd polynomial on the whole half period ------------------------------
// The parity-selected kernels above need, per sine, six per-lane indexed constant loads (LDC), four
// FSEL and a handful of LOP3/SHL for the index and the select -- ~15 issue slots next to 13 FP64
// instructions, on a kernel whose issue port is as loaded as its FP64 pipe.  Reducing only to
// |r| <= 1/2 (in units of pi) leaves a single sign flip: sin(pi u) = (-1)^k sin(pi r), k = rint(u),
// r = u - k exact.  The price is a degree-17 odd polynomial instead of degree 13: +1 FP64
// instruction per sine, but every coefficient is a warp-uniform LDCU shared by all M chains.
// Coefficients: Chebyshev-node interpolation of (sin(pi r)/r - pi)/r^2 on r^2 in [0, 1/4] and of
// (sin r / r - 1)/r^2 on [0, (pi/2)^2], computed with 60-digit arithmetic (tools/fit_trig_poly.py);
// measured error <= 1.5 ulp on 2e5 random arguments.
struct __align__(16) WideTrigConsts {
    double magic;        // 1.5 * 2^52
    double pi_hi, pi_lo; // pi = pi_hi + pi_lo
    double inv_pi;
    double cp[8];        // sin(pi r) = r pi_hi + r (pi_lo + z (cp0 + z cp1 + ... + z^7 cp7)), z = r^2
    double cs[8];        // sin(r)    = r + r z (cs0 + z cs1 + ... + z^7 cs7)
    double cc[8];        // cos(r)    = 1 + z (cc0 + z cc1 + ... + z^7 cc7),  |r| <= pi/2
};

static __constant__ WideTrigConsts kWide = {
    6755399441055744.0,
    3.141592653589793, 1.2246467991473532e-16,
    0.3183098861837907,
    {-5.16771278004997, 2.55016403987734, -0.599264529320343, 0.08214588659675232,
     -0.007370430719634586, 0.00046630087411496363, -2.1906201655131958e-05, 7.725743030789876e-07},
    {-0.16666666666666666, 0.008333333333333316, -0.00019841269841254966, 2.7557319219160833e-06,
     -2.505210761669045e-08, 1.6058977292642743e-10, -7.643969663988074e-13, 2.7314368769379893e-15},
    {-0.5, 0.04166666666666634, -0.0013888888888860694, 2.4801587292441663e-05,
     -2.755731776674132e-07, 2.08766308298722e-09, -1.1464688686901012e-11, 4.6276610380051693e-14}};

// host-visible copy of kWide.cp (capi.cu folds the double gyre's eps into it; the CPU test
// tests/test_trig_poly_cpu.py checks that the two tables are the same numbers)
constexpr double kSinPiCpHost[8] = {
    -5.16771278004997, 2.55016403987734, -0.599264529320343, 0.08214588659675232,
    -0.007370430719634586, 0.00046630087411496363, -2.1906201655131958e-05, 7.725743030789876e-07};

// s[m] = sin(pi u[m]), wide kernel
template <int M>
__device__ __forceinline__ void sinpi_wide_v(const double (&u)[M], double (&s)[M]) {
    bool slow = false;
#pragma unroll
    for (int m = 0; m < M; ++m) slow |= !trig_in_range(u[m]);
    if (slow) {
#pragma unroll
        for (int m = 0; m < M; ++m) s[m] = sincos_slow(3.141592653589793 * u[m]).x;
        return;
    }
    int q[M];
    double r[M], z[M], p[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const double t = u[m] + kWide.magic;
        q[m] = __double2loint(t);
        r[m] = u[m] - (t - kWide.magic);  // exact, |r| <= 1/2
        z[m] = r[m] * r[m];
        p[m] = kWide.cp[7];
    }
#pragma unroll
    for (int k = 6; k >= 0; --k)
#pragma unroll
        for (int m = 0; m < M; ++m) p[m] = fma(p[m], z[m], kWide.cp[k]);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        p[m] = fma(p[m], z[m], kWide.pi_lo);
        s[m] = flip_sign(fma(r[m], kWide.pi_hi, r[m] * p[m]), q[m] & 1);
    }
}

// |u| < 2^50 as an integer compare on the high word (false for NaN / inf): the exact reduction
// k = rint(u), r = u - k of the wide sinpi kernels holds for every such u
__device__ __forceinline__ bool sinpi_in_range(double u) {
    return (unsigned)(__double2hiint(u) & 0x7fffffff) < 0x43100000u;
}

// sin(pi u) in 12 FP64 instructions: the same reduction and polynomial as sinpi_wide_v, with pi
// folded into the polynomial's constant term, sin(pi r) = r (pi + z Q(z)), instead of the split
// r pi_hi + r (pi_lo + z Q(z)) -- two instructions fewer per sine.  Measured on 2e7 random
// arguments (tools/fit_trig_poly.py): max error 2.96 ulp / 3.3e-16 absolute, against 2.33 ulp for
// the split form and 1.4e-15 absolute for the reference's own sin(pi*x) with |x| < 4, whose
// argument pi*x is rounded before libm sees it.
// B200CS_SINPI_NOBRANCH (round 2, default): no libm fall-back branch in the sin(pi u) kernels.  The
// guard is not needed for correctness (see below) and without it a whole step attempt of the
// double-gyre kernel is a handful of basic blocks that ptxas schedules as one: 705 -> 311 non-FP64
// instructions in the attempt loop, 1060 -> 1103 M points/s at 8192^2 (profiles/r2_ab_dg_nobranch.txt).
#ifndef B200CS_SINPI_NOBRANCH
#define B200CS_SINPI_NOBRANCH 1
#endif
template <int M>
__device__ __forceinline__ void sinpi12_v(const double (&u)[M], double (&s)[M]) {
#if !B200CS_SINPI_NOBRANCH
    bool slow = false;
#pragma unroll
    for (int m = 0; m < M; ++m) slow |= !sinpi_in_range(u[m]);
    if (slow) {  // NaN, inf, or |u| >= 2^50 (every such double is an integer: sin(pi u) = +-0)
#pragma unroll
        for (int m = 0; m < M; ++m) s[m] = sincos_slow(3.141592653589793 * u[m]).x;
        return;
    }
#endif
    // (without the guard the fast path is still right at the edges: NaN and inf propagate to NaN
    //  through u - (t - magic); a finite |u| >= 2^52 is an integer and gives r = 0, i.e. +-0)
    int q[M];
    double r[M], z[M], p[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const double t = u[m] + kWide.magic;
        q[m] = __double2loint(t);
        r[m] = u[m] - (t - kWide.magic);  // exact, |r| <= 1/2
#if B200CS_FLIP_EARLY
        r[m] = flip_sign(r[m], q[m] & 1);  // (-1)^k on r: the polynomial is odd, the bits of the result are the same
#endif
        z[m] = r[m] * r[m];
        p[m] = kWide.cp[7];
    }
    // (Estrin's scheme -- depth 4 instead of 7, two more multiplications -- measured slower: 893 vs
    // 998 M points/s at 8192^2, profiles/r1e_ab_estrin.txt: the FP64 pipe, not the chain, is the limit)
#pragma unroll
    for (int k = 6; k >= 0; --k)
#pragma unroll
        for (int m = 0; m < M; ++m) p[m] = fma(p[m], z[m], kWide.cp[k]);
#pragma unroll
    for (int m = 0; m < M; ++m) {
#if B200CS_FLIP_EARLY == 2
        s[m] = __dmul_rn(r[m], fma(p[m], z[m], kWide.pi_hi));   // rounded on its own, as when the flip followed it
#elif B200CS_FLIP_EARLY
        s[m] = r[m] * fma(p[m], z[m], kWide.pi_hi);
#else
        s[m] = flip_sign(r[m] * fma(p[m], z[m], kWide.pi_hi), q[m] & 1);
#endif
    }
}

// amp * sin(pi u) with the amplitude folded into the coefficients: ce = amp * {cp0 .. cp7, pi}
// (formed once per launch on the host).  Same reduction and Horner chain as sinpi12_v; the folded
// coefficients carry one extra rounding each (<= 1 ulp of a coefficient, i.e. the same size as
// the rounding of the product amp * sin it replaces).
template <int M>
__device__ __forceinline__ void sinpi12_scaled_v(const double (&u)[M], double (&s)[M], const double (&ce)[9]) {
#if !B200CS_SINPI_NOBRANCH
    bool slow = false;
#pragma unroll
    for (int m = 0; m < M; ++m) slow |= !sinpi_in_range(u[m]);
    if (slow) {
        const double amp = ce[8] * 0.3183098861837907;   // ce[8] = amp * pi
#pragma unroll
        for (int m = 0; m < M; ++m) s[m] = amp * sincos_slow(3.141592653589793 * u[m]).x;
        return;
    }
#endif
    int q[M];
    double r[M], z[M], p[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const double t = u[m] + kWide.magic;
        q[m] = __double2loint(t);
        r[m] = u[m] - (t - kWide.magic);  // exact, |r| <= 1/2
#if B200CS_FLIP_EARLY
        r[m] = flip_sign(r[m], q[m] & 1);
#endif
        z[m] = r[m] * r[m];
        p[m] = ce[7];
    }
#pragma unroll
    for (int k = 6; k >= 0; --k)
#pragma unroll
        for (int m = 0; m < M; ++m) p[m] = fma(p[m], z[m], ce[k]);
#pragma unroll
    for (int m = 0; m < M; ++m) {
#if B200CS_FLIP_EARLY == 2
        s[m] = __dmul_rn(r[m], fma(p[m], z[m], ce[8]));   // rounded on its own, as when the flip followed it
#elif B200CS_FLIP_EARLY
        s[m] = r[m] * fma(p[m], z[m], ce[8]);
#else
        s[m] = flip_sign(r[m] * fma(p[m], z[m], ce[8]), q[m] & 1);
#endif
    }
}

// (A branch-free form -- no libm fall-back, r forced to 0 for finite |u| >= 2^51 so that a whole step
// attempt is one basic block -- measured 1-2 % slower than the guarded form: profiles/r1c_ab_variants_a.txt,
// variants "nb*" / "w2".)

// s[m] = sin(x[m]), wide kernel (|x| < 1e5, else libm)
template <int M>
__device__ __forceinline__ void sin_wide_v(const double (&x)[M], double (&s)[M]) {
    bool slow = false;
#pragma unroll
    for (int m = 0; m < M; ++m) slow |= !trig_in_range(x[m]);
    if (slow) {
#pragma unroll
        for (int m = 0; m < M; ++m) s[m] = sincos_slow(x[m]).x;
        return;
    }
    int q[M];
    double r[M], z[M], p[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const double t = fma(x[m], kWide.inv_pi, kWide.magic);
        q[m] = __double2loint(t);
        const double k = t - kWide.magic;
        r[m] = fma(-k, kWide.pi_lo, fma(-k, kWide.pi_hi, x[m]));  // two-term Cody-Waite by pi
        z[m] = r[m] * r[m];
        p[m] = kWide.cs[7];
    }
#pragma unroll
    for (int k = 6; k >= 0; --k)
#pragma unroll
        for (int m = 0; m < M; ++m) p[m] = fma(p[m], z[m], kWide.cs[k]);
#pragma unroll
    for (int m = 0; m < M; ++m) s[m] = flip_sign(fma(r[m] * z[m], p[m], r[m]), q[m] & 1);
}

// sin and cos of one argument with the wide kernels: one reduction by pi, two polynomials on
// |r| <= pi/2, one common sign flip -- 22 FP64 instructions and two LOP3 instead of 20 FP64 plus
// the parity selects of sincos_fast.  cos is evaluated as 1 + z Pc(z): its ABSOLUTE error is
// ~2e-16 everywhere, its relative error grows near the zeros of cos (the flows multiply it by an
// O(1) amplitude and add it to O(1) terms, so absolute accuracy is what matters there).
__device__ __forceinline__ void sincos_wide_core(double x, double *s, double *c);
__device__ __forceinline__ void sincos_wide(double x, double *s, double *c) {
    if (!trig_in_range(x)) {
        const double2 sc = sincos_slow(x);
        *s = sc.x;
        *c = sc.y;
        return;
    }
    sincos_wide_core(x, s, c);
}
// the unguarded kernel: the caller has checked |x| < 1e5 (one test for all the arguments of a RHS)
__device__ __forceinline__ void sincos_wide_core(double x, double *s, double *c) {
    const double t = fma(x, kWide.inv_pi, kWide.magic);
    const int q = __double2loint(t);
    const double k = t - kWide.magic;
    const double r = fma(-k, kWide.pi_lo, fma(-k, kWide.pi_hi, x));
    const double z = r * r;
    double ps = kWide.cs[7], pc = kWide.cc[7];
#pragma unroll
    for (int j = 6; j >= 0; --j) {
        ps = fma(ps, z, kWide.cs[j]);
        pc = fma(pc, z, kWide.cc[j]);
    }
    *s = flip_sign(fma(r * z, ps, r), q & 1);
    *c = flip_sign(fma(z, pc, 1.0), q & 1);
}

// expm1(x) for x <= 0 (the Bickley jet's tanh / sech^2 come from em = expm1(-2|Y|)): n = rint(x log2 e),
// r = x - n ln2 (two FMA steps), expm1(r) = r + r^2 Q(r) with a degree-11 Q on |r| <= ln2 / 2, and
// e^x - 1 = 2^n expm1(r) + (2^n - 1) in one FMA (exact 2^n - 1: no cancellation for n <= -1, and
// n = 0 gives expm1(r) itself, so tiny |x| keep full relative accuracy).  The same recipe as CUDA's
// libm expm1, but with the constants in the constant bank (one LDCU each, kept in uniform
// registers) instead of 22 UMOV pairs per call site, no branches, and no overflow side: 1.13 ulp
// on 2e7 arguments in [-64, 0] down to denormals (tests/test_trig_poly_cpu.py).  x below -64 is
// clamped (e^-64 = 1.6e-28 is far below half an ulp of -1).
struct __align__(16) Expm1Consts {
    double log2e, ln2_hi, ln2_lo, pad;
    double q[12];
};
static __constant__ Expm1Consts kExpm1 = {
    1.4426950408889634, 0.6931471805599453, 2.3190468138462996e-17, 0.0,
    {0.5, 0.16666666666666666, 0.04166666666666668, 0.008333333333333333, 0.0013888888888879004,
     0.00019841269841263252, 2.480158733663095e-05, 2.7557319247344507e-06, 2.755726307730941e-07,
     2.5052070959803662e-08, 2.091821231913398e-09, 1.6086677666787362e-10}};

__device__ __forceinline__ double expm1_neg(double x) {
    x = fmax(x, -64.0);
    const double t = fma(x, kExpm1.log2e, kWide.magic);
    const int n = __double2loint(t);
    const double k = t - kWide.magic;
    const double r = fma(-k, kExpm1.ln2_lo, fma(-k, kExpm1.ln2_hi, x));
    double q = kExpm1.q[11];
#pragma unroll
    for (int j = 10; j >= 0; --j) q = fma(q, r, kExpm1.q[j]);
    const double p = fma(r * r, q, r);
    const double s = __hiloint2double((n + 1023) << 20, 0);  // 2^n, n in [-93, 0]
    return fma(p, s, s - 1.0);
}

__device__ __forceinline__ double sin_fast(double x) {
    const double a[1] = {x};
    double s[1];
    sin_v<1>(a, s);
    return s[0];
}

__device__ __forceinline__ double cos_fast(double x) {
    const double a[1] = {x};
    double s[1];
    sin_v<1, 1>(a, s);
    return s[0];
}

// both outputs for one argument (both kernels are needed anyway)
__device__ __forceinline__ void sincos_fast(double x, double *s, double *c) {
    if (!trig_in_range(x)) {
        const double2 sc = sincos_slow(x);
        *s = sc.x;
        *c = sc.y;
        return;
    }
    int q;
    const double r = trig_reduce(x, q);
    const double z = r * r;
    double ps = kTrig.s[5], pc = kTrig.c[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        ps = fma(ps, z, kTrig.s[k]);
        pc = fma(pc, z, kTrig.c[k]);
    }
    const double sr = fma(z * r, ps, r);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    const bool odd = q & 1;
    *s = flip_sign(odd ? cr : sr, (q >> 1) & 1);
    *c = flip_sign(odd ? sr : cr, ((q + 1) >> 1) & 1);
}

}  // namespace b200cs
