// spline.cuh -- device evaluators over get_interp_arrays_2D / get_interp_arrays_scalar outputs.
//
// Replaces interpolation.splines.eval_spline(grid, C, point, k=3, extrap_mode) as called at
// /root/reference/src/numbacs/flows.py:168-253 (velocity, inside the RHS) and flows.py:409
// (scalar vorticity), and eval_linear at flows.py:631 (trilinear scalar).
//
// Layout: coefficients of the two velocity components are interleaved as double2 (u, v) so one
// 16-byte load serves both and the blending weights are computed once per RHS (the reference
// evaluates u and v separately and recomputes them).  C-order (t, x, y), y fastest: the four
// y-taps of one (t, x) pair are one contiguous 64-byte run.
#pragma once
#include <cuda_runtime.h>

#ifndef B200CS_STRICT
#define B200CS_STRICT 0
#endif
#ifndef B200CS_LOCATE_MAGIC
#define B200CS_LOCATE_MAGIC 1
#endif
#ifndef B200CS_STRICT_RHS   // the right-hand-side half of the strict build on its own (A/B decomposition)
#define B200CS_STRICT_RHS B200CS_STRICT
#endif

namespace b200cs {

struct SplineGridDev {
    double a[3];      // axis start
    double b[3];      // axis end
    double delta[3];  // (b - a) / (n - 1)
    double inv_delta[3];
    int n[3];         // data points per axis
    long long s0, s1; // element strides of axes 0 and 1 of the coefficient array
    int extrap;       // B200CS_EXTRAP_*
    unsigned long long *oog;  // velocity fields: counter of evaluations outside the data grid (may be null)
};

// Out-of-grid test on the integer pipe (the FP64 pipe is what these kernels are short of): the bit
// pattern of a local coordinate lam in [+0, 1] is at most 0x3ff0000000000000; anything above 1,
// negative (sign bit set, -0 included) or NaN compares greater as an unsigned 64-bit integer.
__device__ __forceinline__ bool lam_outside(double lam) {
    return (unsigned long long)__double_as_longlong(lam) > 0x3ff0000000000000ULL;
}
__device__ __forceinline__ void count_outside(const SplineGridDev &g, bool outside) {
    if (outside && g.oog) atomicAdd(g.oog, 1ULL);
}

// cell index and local coordinate:  i = clamp(floor((x-a)/delta), 0, n-2),  lam = ((x-a) - i*delta)/delta
// (the product i*delta is rounded before the subtraction, as in unfused CPU arithmetic)
__device__ __forceinline__ void axis_locate(const SplineGridDev &g, int d, double x, int &i, double &lam) {
    const double dd = x - g.a[d];
#if B200CS_STRICT_RHS
    // the oracle's arithmetic: true divisions by delta = (b - a)/(n - 1)
    const double delta = g.delta[d];
    double fi = floor(dd / delta);
    fi = fmin(fmax(fi, 0.0), (double)(g.n[d] - 2));
    i = (int)fi;
    lam = __dsub_rn(dd, __dmul_rn(fi, delta)) / delta;
#elif B200CS_LOCATE_MAGIC
    // floor() and the double -> int conversion without FRND.F64 / F2I.F64 (each costs ~3.7 DFMA issue
    // slots on sm_100, profiles/r2_ubench_fp64_conversions.txt): k = rint(u) by the 1.5 * 2^52 magic
    // constant (its low word is the integer), minus one where the rounding went up; clamped as integers.
    const double u = dd * g.inv_delta[d];
    const double tm = u + 6755399441055744.0;
    const double kr = tm - 6755399441055744.0;
    int k = __double2loint(tm) - (kr > u ? 1 : 0);
    // |u| >= 2^31 (a particle absurdly far outside the grid) or NaN: the clamp decides
    if (!(fabs(u) < 2.0e9)) k = (u > 0.0) ? g.n[d] - 2 : 0;
    k = max(0, min(k, g.n[d] - 2));
    i = k;
    const double fi = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;   // (double)k, k >= 0
    const double r = __dsub_rn(dd, __dmul_rn(fi, g.delta[d]));
    lam = r * g.inv_delta[d];
#else
    double fi = floor(dd * g.inv_delta[d]);
    fi = fmin(fmax(fi, 0.0), (double)(g.n[d] - 2));
    i = (int)fi;
    const double r = __dsub_rn(dd, __dmul_rn(fi, g.delta[d]));
    lam = r * g.inv_delta[d];
#endif
}

// a*b + c: one FMA, or two roundings in the parity-calibration build (dop853.cuh, B200CS_STRICT)
__device__ __forceinline__ double sp_mad(double a, double b, double c) {
#if B200CS_STRICT_RHS
    return __dadd_rn(__dmul_rn(a, b), c);
#else
    return fma(a, b, c);
#endif
}

// uniform cubic B-spline blending weights; outside [0,1] they continue linearly when the
// extrapolation mode is 'linear'.
__device__ __forceinline__ void bspline_weights(double l, bool linear_ext, double (&P)[4]) {
    const double s = 1.0 / 6.0;
#if B200CS_STRICT_RHS
    if (!(linear_ext && (l < 0.0 || l > 1.0))) {
        // the CPU restatement's evaluation order (Horner-free, left to right)
        const double l2 = l * l, l3 = l2 * l;
        P[0] = (-1.0 / 6.0) * l3 + (3.0 / 6.0) * l2 + (-3.0 / 6.0) * l + 1.0 / 6.0;
        P[1] = (3.0 / 6.0) * l3 + (-6.0 / 6.0) * l2 + 4.0 / 6.0;
        P[2] = (-3.0 / 6.0) * l3 + (3.0 / 6.0) * l2 + (3.0 / 6.0) * l + 1.0 / 6.0;
        P[3] = (1.0 / 6.0) * l3;
        return;
    }
#endif
    if (linear_ext && l < 0.0) {
        P[0] = fma(-0.5, l, s);
        P[1] = 4.0 * s;
        P[2] = fma(0.5, l, s);
        P[3] = 0.0;
    } else if (linear_ext && l > 1.0) {
        const double m = l - 1.0;
        P[0] = 0.0;
        P[1] = fma(-0.5, m, s);
        P[2] = 4.0 * s;
        P[3] = fma(0.5, m, s);
    } else {
        const double l2 = l * l, l3 = l2 * l;
        P[0] = fma(-s, l3, fma(0.5, l2, fma(-0.5, l, s)));
        P[1] = fma(0.5, l3, fma(-1.0, l2, 4.0 * s));
        P[2] = fma(-0.5, l3, fma(0.5, l2, fma(0.5, l, s)));
        P[3] = s * l3;
    }
}

// Applies the extrapolation rule to one coordinate. Returns false when the point is outside the
// grid and the mode is 'constant' (value 0).
__device__ __forceinline__ bool extrap_coord(const SplineGridDev &g, int d, double &x) {
    if (g.extrap == B200CS_EXTRAP_CONSTANT) {
        if (x < g.a[d] || x > g.b[d]) return false;
    } else if (g.extrap == B200CS_EXTRAP_NEAREST) {
        x = fmax(g.a[d], fmin(g.b[d], x));
    }
    return true;
}

// cubic weights for a local coordinate inside [0, 1] (no extrapolation cases, no branches)
__device__ __forceinline__ void bspline_weights_in(double l, double (&P)[4]) {
#if B200CS_STRICT_RHS
    bspline_weights(l, false, P);
#else
    const double s = 1.0 / 6.0;
    const double l2 = l * l, l3 = l2 * l;
    P[0] = sp_mad(-s, l3, sp_mad(0.5, l2, sp_mad(-0.5, l, s)));
    P[1] = sp_mad(0.5, l3, sp_mad(-1.0, l2, 4.0 * s));
    P[2] = sp_mad(-0.5, l3, sp_mad(0.5, l2, sp_mad(0.5, l, s)));
    P[3] = s * l3;
#endif
}

// the 64 taps: nested t -> x -> y sum of an interleaved (u, v) field
__device__ __forceinline__ void spline_taps_uv(const SplineGridDev &g, const double2 *__restrict__ C, int i0, int i1,
                                               int i2, const double (&P0)[4], const double (&P1)[4],
                                               const double (&P2)[4], double &u, double &v) {
    const double2 *base = C + (long long)i0 * g.s0 + (long long)i1 * g.s1 + i2;
    double au = 0.0, av = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double bu = 0.0, bv = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double2 *c = base + a * g.s0 + b * g.s1;
            const double2 c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2), c3 = __ldg(c + 3);
            const double cu = sp_mad(P2[3], c3.x, sp_mad(P2[2], c2.x, sp_mad(P2[1], c1.x, P2[0] * c0.x)));
            const double cv = sp_mad(P2[3], c3.y, sp_mad(P2[2], c2.y, sp_mad(P2[1], c1.y, P2[0] * c0.y)));
            bu = sp_mad(P1[b], cu, bu);
            bv = sp_mad(P1[b], cv, bv);
        }
        au = sp_mad(P0[a], bu, au);
        av = sp_mad(P0[a], bv, av);
    }
    u = au;
    v = av;
}

// the general path: a point outside the data grid, every extrapolation mode (rare; out of line so
// that the common path below stays one basic block)
static __device__ __noinline__ double2 eval_spline_uv_outside(const SplineGridDev *gp, const double2 *__restrict__ C,
                                                              double t, double x, double y) {
    const SplineGridDev &g = *gp;
    count_outside(g, true);
    if (!extrap_coord(g, 0, t) || !extrap_coord(g, 1, x) || !extrap_coord(g, 2, y)) return make_double2(0.0, 0.0);
    int i0, i1, i2;
    double l0, l1, l2;
    axis_locate(g, 0, t, i0, l0);
    axis_locate(g, 1, x, i1, l1);
    axis_locate(g, 2, y, i2, l2);
    const bool lin = g.extrap == B200CS_EXTRAP_LINEAR;
    double P0[4], P1[4], P2[4];
    bspline_weights(l0, lin, P0);
    bspline_weights(l1, lin, P1);
    bspline_weights(l2, lin, P2);
    double u, v;
    spline_taps_uv(g, C, i0, i1, i2, P0, P1, P2, u, v);
    return make_double2(u, v);
}

// 64-tap tri-cubic evaluation of an interleaved (u, v) field.  Common path (the point is inside the
// grid, whatever the extrapolation mode): locate, ONE integer test of the three local coordinates,
// cubic weights, taps.  A point outside takes the out-of-line general path above.
__device__ __forceinline__ void eval_spline_uv(const SplineGridDev &g, const double2 *__restrict__ C,
                                               double t, double x, double y, double &u, double &v) {
    int i0, i1, i2;
    double l0, l1, l2;
    axis_locate(g, 0, t, i0, l0);
    axis_locate(g, 1, x, i1, l1);
    axis_locate(g, 2, y, i2, l2);
    // NaN coordinates: lam is NaN, its bit pattern compares greater -> general path -> NaN / 0 as before
    bool outside = lam_outside(l0) || lam_outside(l1) || lam_outside(l2);
    // 'constant' and 'nearest' are defined on the coordinates themselves (a point one ulp beyond the
    // last node is outside even if its local coordinate rounds to 1): the exact comparisons, taken
    // only in those modes ('linear' continues the cubic weights, so lam decides)
    if (g.extrap != B200CS_EXTRAP_LINEAR)
        outside = outside || t < g.a[0] || t > g.b[0] || x < g.a[1] || x > g.b[1] || y < g.a[2] || y > g.b[2];
    if (outside) {
        const double2 r = eval_spline_uv_outside(&g, C, t, x, y);
        u = r.x;
        v = r.y;
        return;
    }
    double P0[4], P1[4], P2[4];
    bspline_weights_in(l0, P0);
    bspline_weights_in(l1, P1);
    bspline_weights_in(l2, P2);
    spline_taps_uv(g, C, i0, i1, i2, P0, P1, P2, u, v);
}

// scalar tri-cubic
__device__ __forceinline__ double eval_spline_s(const SplineGridDev &g, const double *__restrict__ C,
                                                double t, double x, double y) {
    if (!extrap_coord(g, 0, t) || !extrap_coord(g, 1, x) || !extrap_coord(g, 2, y)) return 0.0;
    int i0, i1, i2;
    double l0, l1, l2;
    axis_locate(g, 0, t, i0, l0);
    axis_locate(g, 1, x, i1, l1);
    axis_locate(g, 2, y, i2, l2);
    const bool lin = g.extrap == B200CS_EXTRAP_LINEAR;
    double P0[4], P1[4], P2[4];
    bspline_weights(l0, lin, P0);
    bspline_weights(l1, lin, P1);
    bspline_weights(l2, lin, P2);
    const double *base = C + (long long)i0 * g.s0 + (long long)i1 * g.s1 + i2;
    double acc0 = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double acc1 = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double *c = base + a * g.s0 + b * g.s1;
            const double acc2 =
                sp_mad(P2[3], __ldg(c + 3), sp_mad(P2[2], __ldg(c + 2), sp_mad(P2[1], __ldg(c + 1), P2[0] * __ldg(c))));
            acc1 = sp_mad(P1[b], acc2, acc1);
        }
        acc0 = sp_mad(P0[a], acc1, acc0);
    }
    return acc0;
}

// trilinear evaluation of an interleaved (u, v) field on the raw (n0, n1, n2) data: eval_linear at
// flows.py:470-503 (get_flow_linear_2D), one 16-byte load per tap
__device__ __forceinline__ void eval_linear_uv(const SplineGridDev &g, const double2 *__restrict__ F,
                                               double t, double x, double y, double &u, double &v) {
    u = 0.0;
    v = 0.0;
    const double t_in = t, x_in = x, y_in = y;
    if (!extrap_coord(g, 0, t) || !extrap_coord(g, 1, x) || !extrap_coord(g, 2, y)) {
        count_outside(g, true);
        return;
    }
    int i0, i1, i2;
    double l0, l1, l2;
    axis_locate(g, 0, t, i0, l0);
    axis_locate(g, 1, x, i1, l1);
    axis_locate(g, 2, y, i2, l2);
    if (g.extrap == B200CS_EXTRAP_NEAREST) count_outside(g, t != t_in || x != x_in || y != y_in);   // clamped
    else count_outside(g, lam_outside(l0) || lam_outside(l1) || lam_outside(l2));
    const double2 *c = F + (long long)i0 * g.s0 + (long long)i1 * g.s1 + i2;
    const double m2 = 1.0 - l2;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double wa = a ? l0 : 1.0 - l0;
        double vu = 0.0, vv = 0.0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double wb = b ? l1 : 1.0 - l1;
            const double2 *cc = c + a * g.s0 + b * g.s1;
            const double2 c0 = __ldg(cc), c1 = __ldg(cc + 1);
            vu = sp_mad(wb, sp_mad(l2, c1.x, m2 * c0.x), vu);
            vv = sp_mad(wb, sp_mad(l2, c1.y, m2 * c0.y), vv);
        }
        u = sp_mad(wa, vu, u);
        v = sp_mad(wa, vv, v);
    }
}

// scalar trilinear on the raw (n0, n1, n2) data
__device__ __forceinline__ double eval_linear_s(const SplineGridDev &g, const double *__restrict__ F,
                                                double t, double x, double y) {
    if (!extrap_coord(g, 0, t) || !extrap_coord(g, 1, x) || !extrap_coord(g, 2, y)) return 0.0;
    int i0, i1, i2;
    double l0, l1, l2;
    axis_locate(g, 0, t, i0, l0);
    axis_locate(g, 1, x, i1, l1);
    axis_locate(g, 2, y, i2, l2);
    const double *c = F + (long long)i0 * g.s0 + (long long)i1 * g.s1 + i2;
    double v = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double wa = a ? l0 : 1.0 - l0;
        double va = 0.0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double wb = b ? l1 : 1.0 - l1;
            const double *cc = c + a * g.s0 + b * g.s1;
            va = sp_mad(wb, sp_mad(l2, __ldg(cc + 1), (1.0 - l2) * __ldg(cc)), va);
        }
        v = sp_mad(wa, va, v);
    }
    return v;
}

// The same two evaluators on a TIME-COLLAPSED slab W[m, n] = sum_a P0[a] C[i0 + a, m, n] (cubic) or
// (1 - l0) F[i0] + l0 F[i0 + 1] (trilinear): LAVD evaluates the vorticity at the n output times
// t_k only, the same for every particle, so the time contraction is done once per output time
// (vort_slab_kernel, diag_kernels.cu) and each evaluation gathers 16 taps (4) instead of 64 (8).
// The sum is the reference's nested t -> x -> y sum re-associated (t first): a rounding-level change.
__device__ __forceinline__ double eval_spline_s2(const SplineGridDev &g, const double *__restrict__ W, double x,
                                                 double y) {
    if (!extrap_coord(g, 1, x) || !extrap_coord(g, 2, y)) return 0.0;
    int i1, i2;
    double l1, l2;
    axis_locate(g, 1, x, i1, l1);
    axis_locate(g, 2, y, i2, l2);
    const bool lin = g.extrap == B200CS_EXTRAP_LINEAR;
    double P1[4], P2[4];
    bspline_weights(l1, lin, P1);
    bspline_weights(l2, lin, P2);
    const double *base = W + (long long)i1 * g.s1 + i2;
    double acc1 = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double *c = base + b * g.s1;
        const double acc2 =
            sp_mad(P2[3], __ldg(c + 3), sp_mad(P2[2], __ldg(c + 2), sp_mad(P2[1], __ldg(c + 1), P2[0] * __ldg(c))));
        acc1 = sp_mad(P1[b], acc2, acc1);
    }
    return acc1;
}

__device__ __forceinline__ double eval_linear_s2(const SplineGridDev &g, const double *__restrict__ W, double x,
                                                 double y) {
    if (!extrap_coord(g, 1, x) || !extrap_coord(g, 2, y)) return 0.0;
    int i1, i2;
    double l1, l2;
    axis_locate(g, 1, x, i1, l1);
    axis_locate(g, 2, y, i2, l2);
    const double *c = W + (long long)i1 * g.s1 + i2;
    double va = 0.0;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const double wb = b ? l1 : 1.0 - l1;
        const double *cc = c + b * g.s1;
        va = sp_mad(wb, sp_mad(l2, __ldg(cc + 1), (1.0 - l2) * __ldg(cc)), va);
    }
    return va;
}

// a scalar field (vorticity): cubic-spline coefficients or the raw field for trilinear evaluation;
// W (optional): slabs collapsed at the output times, slab k at W + k * wstride (layout of one time
// level of C: g.s1 elements per row)
struct ScalarDev {
    SplineGridDev g;
    const double *C;
    int linear;
    const double *W;
    long long wstride;
};

__device__ __forceinline__ double scalar_at(const ScalarDev &S, double t, double x, double y) {
    return S.linear ? eval_linear_s(S.g, S.C, t, x, y) : eval_spline_s(S.g, S.C, t, x, y);
}
// the field at output time k (t == tspan[k]): from the collapsed slab when there is one
__device__ __forceinline__ double scalar_at_k(const ScalarDev &S, long long k, double t, double x, double y) {
    if (S.W == nullptr) return scalar_at(S, t, x, y);
    const double *Wk = S.W + k * S.wstride;
    return S.linear ? eval_linear_s2(S.g, Wk, x, y) : eval_spline_s2(S.g, Wk, x, y);
}

// Python float modulo (result takes the sign of the divisor), diagnostics.py:350-376
__device__ __forceinline__ double pymod_any(double a, double m) {
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}

}  // namespace b200cs
