// fastmath.cuh -- lean FP64 sin/cos for the right-hand sides.
//
// The RHS of the analytic flows is dominated by sin/cos (flows.py:1152-1158: five per call).
// CUDA's libm versions are accurate but wrap ~24 FP64 instructions in ~60 integer / uniform /
// branch instructions (immediates materialised with UMOV pairs, inf/NaN checks, a Payne-Hanek
// slow path behind a CALL).  On B200 the FP64 pipe issues one warp instruction every two cycles
// per SM sub-partition, so that overhead -- not FP64 throughput -- would bound the kernel.  These
// versions keep the same numerical recipe (3-term Cody-Waite reduction by pi/2 with FMA, the
// classical minimax kernels on [-pi/4, pi/4], error < 1 ulp) but read every constant straight from
// the constant bank as an instruction operand and use a branch-free quadrant fix-up.  Arguments
// with |x| >= 1e5 (never seen by the flows' own domains) fall back to libm.
#pragma once
#include <cuda_runtime.h>

namespace b200cs {

struct __align__(16) TrigConsts {
    double two_over_pi, magic, p1, p2, p3;
    double s[6];  // sin kernel: x + x^3 (s0 + z s1 + ... + z^5 s5)
    double c[6];  // cos kernel: 1 - z/2 + z^2 (c0 + z c1 + ... + z^5 c5)
};

static __constant__ TrigConsts kTrig = {
    0.63661977236758138,      // 2/pi
    6755399441055744.0,       // 1.5 * 2^52: adding it rounds to nearest integer
    1.5707963267948966,       // pi/2 = p1 + p2 + p3 (+ O(1e-48)); p1 = fl(pi/2)
    6.123233995736757e-17,    // 0x3c91a62633145c00: trailing zero bits keep k*p2 exact for |k| < 2^16
    8.478427660368898e-32,    // 0x397b839a252049c0
    {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
     2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10},
    {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
     -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11}};

// libm fall-backs for huge / non-finite arguments, kept out of line so that their Payne-Hanek
// code is not replicated at every call site
static __device__ __noinline__ double2 sincos_slow(double x) {
    double2 r;
    sincos(x, &r.x, &r.y);
    return r;
}
static __device__ __noinline__ double sin_slow(double x) { return sin(x); }
static __device__ __noinline__ double cos_slow(double x) { return cos(x); }

// r = x - k*pi/2 with k = rint(x * 2/pi); returns k's low bits in q.  |x| < 1e5.
__device__ __forceinline__ double trig_reduce(double x, int &q) {
    const double t = fma(x, kTrig.two_over_pi, kTrig.magic);
    q = __double2loint(t);
    const double k = t - kTrig.magic;
    double r = fma(-k, kTrig.p1, x);
    r = fma(-k, kTrig.p2, r);
    r = fma(-k, kTrig.p3, r);
    return r;
}

__device__ __forceinline__ double sin_kernel(double r, double z) {
    double p = kTrig.s[5];
    p = fma(p, z, kTrig.s[4]);
    p = fma(p, z, kTrig.s[3]);
    p = fma(p, z, kTrig.s[2]);
    p = fma(p, z, kTrig.s[1]);
    p = fma(p, z, kTrig.s[0]);
    return fma(z * r, p, r);
}

__device__ __forceinline__ double cos_kernel(double z) {
    double p = kTrig.c[5];
    p = fma(p, z, kTrig.c[4]);
    p = fma(p, z, kTrig.c[3]);
    p = fma(p, z, kTrig.c[2]);
    p = fma(p, z, kTrig.c[1]);
    p = fma(p, z, kTrig.c[0]);
    return fma(z * z, p, fma(z, -0.5, 1.0));
}

// flips the sign of d when bit is non-zero (bit is 0 or 1)
__device__ __forceinline__ double flip_sign(double d, int bit) {
    return __hiloint2double(__double2hiint(d) ^ (bit << 31), __double2loint(d));
}

__device__ __forceinline__ void sincos_fast(double x, double *s, double *c) {
    if (!(fabs(x) < 1.0e5)) {  // also catches NaN / inf
        const double2 sc = sincos_slow(x);
        *s = sc.x;
        *c = sc.y;
        return;
    }
    int q;
    const double r = trig_reduce(x, q);
    const double z = r * r;
    const double sr = sin_kernel(r, z), cr = cos_kernel(z);
    const bool odd = q & 1;
    const double ss = odd ? cr : sr;
    const double cc = odd ? sr : cr;
    *s = flip_sign(ss, (q >> 1) & 1);
    *c = flip_sign(cc, ((q + 1) >> 1) & 1);
}

__device__ __forceinline__ double sin_fast(double x) {
    if (!(fabs(x) < 1.0e5)) return sin_slow(x);
    int q;
    const double r = trig_reduce(x, q);
    const double z = r * r;
    const double v = (q & 1) ? cos_kernel(z) : sin_kernel(r, z);
    return flip_sign(v, (q >> 1) & 1);
}

__device__ __forceinline__ double cos_fast(double x) {
    if (!(fabs(x) < 1.0e5)) return cos_slow(x);
    int q;
    const double r = trig_reduce(x, q);
    const double z = r * r;
    const double v = (q & 1) ? sin_kernel(r, z) : cos_kernel(z);
    return flip_sign(v, ((q + 1) >> 1) & 1);
}

}  // namespace b200cs
