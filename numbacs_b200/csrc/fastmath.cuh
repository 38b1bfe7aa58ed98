// fastmath.cuh -- lean FP64 sin / cos for the right-hand sides.
//
// The RHS of the analytic flows is dominated by sin/cos (flows.py:1152-1158: five per call).
// CUDA's libm versions are accurate but wrap ~24 FP64 instructions in ~60 integer / uniform /
// branch instructions (immediates materialised with UMOV pairs, inf/NaN checks, a Payne-Hanek
// slow path behind a CALL).  On B200 the FP64 pipe issues one warp instruction every two cycles
// per SM sub-partition and an FP64 instruction cannot take a constant-bank operand directly, so
// that overhead -- not FP64 throughput -- would bound the kernel.  These versions keep the same
// numerical recipe (3-term Cody-Waite reduction by pi/2 with FMA, the classical minimax kernels
// on [-pi/4, pi/4], error < 1 ulp) with a branch-free quadrant fix-up, and evaluate M independent
// arguments side by side (M dependency chains, each constant fetched once).  Arguments with
// |x| >= 1e5 (never produced by the flows' own domains) fall back to libm, out of line.
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

// flips the sign of d when bit is 1
__device__ __forceinline__ double flip_sign(double d, int bit) {
    return __hiloint2double(__double2hiint(d) ^ (bit << 31), __double2loint(d));
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
