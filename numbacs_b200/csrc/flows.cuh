// flows.cuh -- device right-hand sides: the GPU-side flow registry behind `funcptr`.
//
// Each functor is the device implementation of one numba @cfunc(lsoda_sig) the reference builds
// (signature rhs(t, y, dy, p), /root/reference/src/numbacs/flows.py):
//   DoubleGyre  flows.py:1146-1158      BickleyJet  flows.py:1182-1213
//   Abc         flows.py:1249-1258      Spline2D    flows.py:156-253 (spherical 0 / 1 / 2)
//   Spline2D<S, true>  flows.py:458-503 (get_flow_linear_2D)
// Conventions kept from the reference: tt = p[0]*t, dy = p[0]*v(tt, y) (userguide.rst:217-227);
// `params` is used verbatim; operation order follows the reference expressions (FMA contraction
// is allowed, it perturbs results at the 1-ulp level only).
//
// Interface used by the integrator:
//   kAux                      1 if the RHS has a part that depends on time only, else 0
//   time_part<M>(t[M], aux[M]) evaluates that part at M times at once (M independent chains:
//                             the integrator batches the stage times of a step, which are known
//                             before any stage is evaluated)
//   eval(aux, t, y, dy)       the RHS proper
#pragma once
#include <cuda_runtime.h>

#include "fastmath.cuh"
#include "spline.cuh"

namespace b200cs {

constexpr int kMaxParams = 16;
constexpr double kPi = 3.141592653589793;

struct RhsParams {
    double p[kMaxParams];
    SplineGridDev grid;       // Spline2D only
    const double2 *coef_uv;   // Spline2D only
    double r;                 // Spline2D spherical radius
    double d[8];              // constants derived from p on the host (fill_rhs, capi.cu)
    double e[9];              // double gyre: eps * (sinpi polynomial coefficients cp[0..7], pi)
    double e2[9];             // double gyre: (p0 pi A / 2) * (the same): the amplitude of S+ and S-
    unsigned long long *oog;  // Spline2D: out-of-grid evaluation counter of the flow (may be null)
    int slot;                 // >= 0: a copy of this struct sits in c_rhs_slots[slot] (out-of-line RHS)
};

// Parameter blocks for the out-of-line right-hand sides.  A __noinline__ device function cannot
// name its caller's kernel parameters, and reading them through a generic pointer turns ~25
// constant-bank operands per spline RHS into LD.E loads on the same LSU pipe that carries the 64
// coefficient taps.  The launcher copies the block into a __constant__ slot (one slot per stream,
// flowmap_kernel.cuh) and the out-of-line function indexes it: uniform constant loads again.
constexpr int kRhsSlots = 32;
static __constant__ RhsParams c_rhs_slots[kRhsSlots];

#ifndef B200CS_BICKLEY_WIDE
#define B200CS_BICKLEY_WIDE 1
#endif
#ifndef B200CS_ABC_WIDE
#define B200CS_ABC_WIDE 1
#endif
#ifndef B200CS_DG_TRIM      // round 2: eps folded into the sinpi coefficients of a(t), f / df without 2a
#define B200CS_DG_TRIM 1
#endif
#ifndef B200CS_SPLINE_LEAN_SPH   // spherical tail of the spline RHS without fmod / IEEE-division slow paths
#define B200CS_SPLINE_LEAN_SPH 1
#endif
#ifndef B200CS_SPLINE_NOINLINE
#define B200CS_SPLINE_NOINLINE 1
#endif
#ifndef B200CS_BICKLEY_ONEGUARD   // one range test per RHS instead of one per sincos: 212 -> 229 M points/s (config 2)
#define B200CS_BICKLEY_ONEGUARD 1
#endif
#ifndef B200CS_BICKLEY_RCP
#define B200CS_BICKLEY_RCP 1   // 229 -> 249 M points/s (config 2), parity figures unchanged
#endif
#ifndef B200CS_BICKLEY_FOLD   // products of launch constants formed once on the host (fill_rhs): ~10 DMUL fewer per RHS
#define B200CS_BICKLEY_FOLD 1
#endif
#ifndef B200CS_BICKLEY_NOINLINE
#define B200CS_BICKLEY_NOINLINE 0
#endif
// the folded form is what the kernel evaluates AND what fill_rhs (capi.cu) prepares constants for
#define B200CS_BICKLEY_FOLDED \
    (B200CS_BICKLEY_WIDE && B200CS_BICKLEY_ONEGUARD && B200CS_BICKLEY_RCP && B200CS_BICKLEY_FOLD && !B200CS_STRICT_RHS)
#ifndef B200CS_LEAN_F
#define B200CS_LEAN_F 1
#endif
#ifndef B200CS_SINPI_WIDE
#define B200CS_SINPI_WIDE 3
#endif
#ifndef B200CS_TIME_TURNS
#define B200CS_TIME_TURNS 1
#endif
#ifndef B200CS_SIN_WIDE
#define B200CS_SIN_WIDE 1
#endif

// DAMPED = false is the flow with alpha == 0 (the default parameters, and every example of the
// reference): the launcher picks the instantiation, so the hot loop carries neither the damping
// terms nor a test for them.
#if B200CS_STRICT_RHS
// Parity-calibration build: the reference's expressions (flows.py:1152-1158) token for token, every
// operation rounded separately (-fmad=false), CUDA libm sin / cos.
template <bool DAMPED>
struct DoubleGyreT {
    static constexpr int N = 2;
    static constexpr int kAux = 0;
    const RhsParams &P;
    __device__ __forceinline__ explicit DoubleGyreT(const RhsParams &P_) : P(P_) {}
    template <int M>
    __device__ __forceinline__ void time_part(const double (&)[M], double (&)[M]) const {}
    __device__ __forceinline__ void eval(double, double t, const double (&y)[2], double (&dy)[2]) const {
        const double *p = P.p;
        const double pi = kPi;
        const double tt = p[0] * t;
        const double a = p[2] * sin(p[4] * tt + p[5]);
        const double b = 1 - 2 * a;
        const double f = a * (y[0] * y[0]) + b * y[0];
        const double df = 2 * a * y[0] + b;
        dy[0] = p[0] * (-pi * p[1] * sin(pi * f) * cos(pi * y[1]) - p[3] * y[0]);
        dy[1] = p[0] * (pi * p[1] * cos(pi * f) * sin(pi * y[1]) * df - p[3] * y[1]);
    }
};
#else
template <bool DAMPED>
struct DoubleGyreT {
    static constexpr int N = 2;
    static constexpr int kAux = 1;
    static constexpr bool kAuxAffine = B200CS_TIME_TURNS != 0;
    const RhsParams &P;
    __device__ __forceinline__ explicit DoubleGyreT(const RhsParams &P_) : P(P_) {}

    // a(t) = eps * sin(omega*tt + psi), tt = p0*t   (flows.py:1152-1153).
    // p[0] is the integration direction, +-1 (userguide.rst:217-227), so folding it into omega
    // (and into the amplitudes below) is an exact sign change, not a rounding.  The folded
    // constants are formed once on the host (derive_rhs in capi.cu): P.d[0] = omega*p0,
    // P.d[1] = p0*pi*A/2, P.d[2] = -p0*alpha -- one LDCU each instead of being re-derived from
    // the constant bank in every stage.
    // With B200CS_TIME_TURNS the phase is kept in half-turns, u = (omega p0/pi) t + psi/pi
    // (P.d[3], P.d[4] from the host), so that a(t) = eps sin(pi u) uses the exact reduction of the
    // sinpi kernel (no Cody-Waite steps: one FP64 instruction fewer per stage time); and the
    // phase at the stage times x + c_s h is affine in c_s, u_s = c_s (wq h) + (wq x + psiq), so
    // the stage times themselves are never formed.  The phase carries the same kind of rounding
    // (<= ~1 ulp of the phase) as the reference's omega*tt + psi.
    template <int M>
    __device__ __forceinline__ void time_part_affine(double x, double h, const double (&c)[M],
                                                     double (&aux)[M]) const {
        const double eps = P.p[2];
        const double theta = P.d[3] * h, phi = fma(P.d[3], x, P.d[4]);
        double u[M], sa[M];
#pragma unroll
        for (int m = 0; m < M; ++m) u[m] = fma(c[m], theta, phi);
#if B200CS_DG_TRIM
        // a(t) = eps sin(pi u) with eps folded into the polynomial's coefficients on the host
        // (P.e = eps * {cp0..cp7, pi}, fill_rhs): one multiplication fewer per stage time
        (void)eps;
        sinpi12_scaled_v<M>(u, aux, P.e);
        (void)sa;
#else
        sinpi12_v<M>(u, sa);
#pragma unroll
        for (int m = 0; m < M; ++m) aux[m] = eps * sa[m];
#endif
    }

    template <int M>
    __device__ __forceinline__ void time_part(const double (&t)[M], double (&aux)[M]) const {
        const double eps = P.p[2];
        double arg[M], sa[M];
#if B200CS_TIME_TURNS && B200CS_DG_TRIM
#pragma unroll
        for (int m = 0; m < M; ++m) arg[m] = fma(P.d[3], t[m], P.d[4]);
        (void)eps;
        (void)sa;
        sinpi12_scaled_v<M>(arg, aux, P.e);
        return;
#elif B200CS_TIME_TURNS
#pragma unroll
        for (int m = 0; m < M; ++m) arg[m] = fma(P.d[3], t[m], P.d[4]);
        sinpi12_v<M>(arg, sa);
#else
#pragma unroll
        for (int m = 0; m < M; ++m) arg[m] = fma(P.d[0], t[m], P.p[5]);
#if B200CS_SIN_WIDE
        sin_wide_v<M>(arg, sa);
#else
        sin_v<M>(arg, sa);
#endif
#endif
#pragma unroll
        for (int m = 0; m < M; ++m) aux[m] = eps * sa[m];
    }

    // The reference evaluates  -pi A sin(pi f) cos(pi y)  and  pi A cos(pi f) sin(pi y) df
    // (flows.py:1157-1158): two sin and two cos.  With S+ = sin(pi (f + y)), S- = sin(pi (f - y))
    //   sin(pi f) cos(pi y) = (S+ + S-)/2,     cos(pi f) sin(pi y) = (S+ - S-)/2
    // so two sines do the work of two sincos pairs: 16 fewer FP64 instructions per RHS (-21 %) on
    // a kernel that is FP64-issue bound.  The identity is exact; the rounding differs from the
    // reference's product form by a few 1e-16 in ABSOLUTE terms (the same size as libm-vs-libm
    // differences), both forms vanish exactly on the walls x = 0 and y = 0, and the parity tests
    // (step-count equality, 1e-8 x domain) are the guard.  (Forming S+ + S- and S+ - S- as FMAs on
    // r1 Q1 would save one more instruction but leaves the rounding residual of r Q on the walls,
    // where the reference's velocity component is an exact 0: not done.)
    __device__ __forceinline__ void eval(double a, double /*t*/, const double (&y)[2], double (&dy)[2]) const {
        const double c = P.d[1];     // p0 * pi*A/2 (exact scalings of pi*A)
        const double b = 1.0 - 2.0 * a;
#if B200CS_DG_TRIM
        const double g = fma(a, y[0], b);                // a*y0 + b
        const double f = y[0] * g;                       // a*y0**2 + b*y0
        const double df = fma(a, y[0], g);               // 2*a*y0 + b without forming 2a
#elif B200CS_LEAN_F
        const double f = y[0] * fma(a, y[0], b);         // a*y0**2 + b*y0, one instruction fewer
        const double df = fma(2.0 * a, y[0], b);
#else
        const double f = fma(a, y[0] * y[0], b * y[0]);  // a*y0**2 + b*y0
        const double df = fma(2.0 * a, y[0], b);
#endif
        double s[2];
#if B200CS_DG_TRIM && B200CS_SINPI_WIDE == 3 && !defined(B200CS_DG_NO_SINPI)
        // the amplitude c = p0 pi A / 2 is folded into the polynomial of S+ and S- on the host (P.e2), like
        // eps into a(t): two multiplications fewer per RHS.  Both wall identities survive: at x = 0 the
        // arguments are +-y (r, z and the polynomial agree, the signs are opposite: the sum is an exact 0),
        // at y = 0 they coincide (the difference is an exact 0).
        const double arg[2] = {f + y[1], f - y[1]};
        sinpi12_scaled_v<2>(arg, s, P.e2);
        (void)c;
        if (DAMPED) {
            const double damp = P.d[2];  // -p0*alpha
            dy[0] = fma(damp, y[0], -(s[0] + s[1]));
            dy[1] = fma(s[0] - s[1], df, damp * y[1]);
        } else {
            dy[0] = -(s[0] + s[1]);
            dy[1] = (s[0] - s[1]) * df;
        }
#else
#ifdef B200CS_DG_NO_SINPI
        const double arg[2] = {kPi * (f + y[1]), kPi * (f - y[1])};
        sin_v<2>(arg, s);
#else
        const double arg[2] = {f + y[1], f - y[1]};
#if B200CS_SINPI_WIDE == 3
        sinpi12_v<2>(arg, s);
#elif B200CS_SINPI_WIDE
        sinpi_wide_v<2>(arg, s);
#else
        sinpi_v<2>(arg, s);  // sin(pi (f +- y)) with an exact argument reduction
#endif
#endif
        if (DAMPED) {
            const double damp = P.d[2];  // -p0*alpha
            dy[0] = fma(-c, s[0] + s[1], damp * y[0]);
            dy[1] = fma(c * (s[0] - s[1]), df, damp * y[1]);
        } else {
            dy[0] = -c * (s[0] + s[1]);
            dy[1] = (c * (s[0] - s[1])) * df;
        }
#endif
    }
};
#endif  // B200CS_STRICT_RHS
using DoubleGyre = DoubleGyreT<false>;
using DoubleGyreDamped = DoubleGyreT<true>;

struct BickleyJet {
    static constexpr int N = 2;
    static constexpr int kAux = 0;
    static constexpr bool kOutOfLine = B200CS_BICKLEY_NOINLINE != 0;
    const RhsParams &P;
    __device__ __forceinline__ explicit BickleyJet(const RhsParams &P_) : P(P_) {}
    template <int M>
    __device__ __forceinline__ void time_part(const double (&)[M], double (&)[M]) const {}
#if B200CS_BICKLEY_NOINLINE
    // The RHS is ~250 instructions; inlined into the 12 stages + FSAL + hinit it makes the attempt
    // loop 53 KB, well above the 32 KB instruction cache (ncu: no_instruction 2.4 cycles per issue).
    // Out of line it is ONE copy that every stage calls: the loop drops below the cache size and the
    // call costs ~5 % of the body.
    static __device__ __noinline__ double2 eval_ool(const RhsParams *Pp, double t, double y0, double y1) {
        const BickleyJet self(*Pp);
        const double y[2] = {y0, y1};
        double dy[2];
        self.eval_body(t, y, dy);
        return make_double2(dy[0], dy[1]);
    }
    static __device__ __noinline__ double2 eval_ool_slot(int slot, double t, double y0, double y1) {
        const BickleyJet self(c_rhs_slots[slot]);
        const double y[2] = {y0, y1};
        double dy[2];
        self.eval_body(t, y, dy);
        return make_double2(dy[0], dy[1]);
    }
    __device__ __forceinline__ void eval(double, double t, const double (&y)[2], double (&dy)[2]) const {
        const double2 r = (P.slot >= 0) ? eval_ool_slot(P.slot, t, y[0], y[1]) : eval_ool(&P, t, y[0], y[1]);
        dy[0] = r.x;
        dy[1] = r.y;
    }
#else
    __device__ __forceinline__ void eval(double, double t, const double (&y)[2], double (&dy)[2]) const {
        eval_body(t, y, dy);
    }
#endif
    __device__ __forceinline__ void eval_body(double t, const double (&y)[2], double (&dy)[2]) const {
        const double *p = P.p;
        const double tt = p[0] * t;
#if B200CS_STRICT_RHS
        {   // flows.py:1189-1213 token for token (cosh(Y) ** 2 is cosh(Y) * cosh(Y) in numba)
            const double Y = y[1] / p[2];
            const double ch = cosh(Y);
            const double sech2 = 1 / (ch * ch);
            dy[0] = p[0] * (p[1] * sech2 +
                            2 * p[1] * tanh(Y) * sech2 *
                                (p[3] * cos(p[6] * (y[0] - p[9] * tt)) + p[4] * cos(p[7] * (y[0] - p[10] * tt)) +
                                 p[5] * cos(p[8] * (y[0] - p[11] * tt))));
            dy[1] = -p[0] * (p[1] * p[2] * sech2 *
                             (p[3] * p[6] * sin(p[6] * (y[0] - p[9] * tt)) +
                              p[4] * p[7] * sin(p[7] * (y[0] - p[10] * tt)) +
                              p[5] * p[8] * sin(p[8] * (y[0] - p[11] * tt))));
            return;
        }
#endif
#if B200CS_BICKLEY_FOLDED
        {   // The same formulas with every product of launch constants taken from the host (fill_rhs,
            // capi.cu; p0 = +-1 is the integration direction, so folding it is an exact sign change):
            //   e[0..2] = -p0 c_n            a_n = k_n (y0 + e_n t)
            //   e[3..5] = 2 A_n              csum2 = sum 2 A_n cos a_n
            //   e[6..8] = -L A_n k_n         ssum  = sum -L A_n k_n sin a_n
            //   d[6] = 4 p0 U0, d[7] = L / 2, d[5] = 2 / L
            //   S = p0 U0 sech^2 Y,  dy0 = S (1 + tanh Y csum2),  dy1 = S ssum
            // 125 instead of 135 FP64 instructions per evaluation; each folded product is rounded once
            // instead of being re-associated per call, a 1-ulp-level change like the FMA contraction.
            const double *e = P.e;
            const double q0 = y[1] * P.d[5];
            const double Y2 = fma(fma(-q0, P.d[7], y[1]), P.d[5], q0);   // 2 Y = y1 / (L / 2), correctly rounded
            const double em = expm1_neg(-fabs(Y2));
            const double den = 2.0 + em;
            double inv;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(den));
            inv = fma(inv, fma(-den, inv, 1.0), inv);
            inv = fma(inv, fma(-den, inv, 1.0), inv);
            const double th = copysign(-em * inv, Y2);
            const double S = (P.d[6] * (1.0 + em)) * (inv * inv);
            const double a1 = p[6] * fma(e[0], t, y[0]), a2 = p[7] * fma(e[1], t, y[0]), a3 = p[8] * fma(e[2], t, y[0]);
            double s1, c1, s2, c2, s3, c3;
            if (trig_in_range(a1) && trig_in_range(a2) && trig_in_range(a3)) {
                sincos_wide_core(a1, &s1, &c1);
                sincos_wide_core(a2, &s2, &c2);
                sincos_wide_core(a3, &s3, &c3);
            } else {
                const double2 q1 = sincos_slow(a1), q2 = sincos_slow(a2), q3 = sincos_slow(a3);
                s1 = q1.x; c1 = q1.y; s2 = q2.x; c2 = q2.y; s3 = q3.x; c3 = q3.y;
            }
            const double csum2 = fma(e[5], c3, fma(e[4], c2, e[3] * c1));
            const double ssum = fma(e[8], s3, fma(e[7], s2, e[6] * s1));
            dy[0] = S * fma(th, csum2, 1.0);
            dy[1] = S * ssum;
            return;
        }
#endif
#if B200CS_BICKLEY_WIDE
        // y1 / L_y with the reciprocal from the host and one FMA residual correction: the correctly
        // rounded quotient in three instructions (the divisor is a launch constant)
        const double q0 = y[1] * P.d[5];
        const double Y = fma(fma(-q0, p[2], y[1]), P.d[5], q0);
#else
        const double Y = y[1] / p[2];
#endif
        // sech^2(Y) and tanh(Y) from ONE expm1 and one division instead of libm cosh + tanh + a
        // division (the reference's 1/cosh(Y)**2 and tanh(Y), flows.py:1190-1200): with
        // em = expm1(-2|Y|),  tanh|Y| = -em/(2 + em),  sech^2 = 4 (1 + em)/(2 + em)^2; no
        // cancellation anywhere (em is accurate near 0, 2 + em is in [1, 2]).
#if B200CS_BICKLEY_WIDE
        const double em = expm1_neg(-2.0 * fabs(Y));
#else
        const double em = expm1(-2.0 * fabs(Y));
#endif
#if B200CS_BICKLEY_RCP
        // 2 + em lies in (1, 2]: MUFU reciprocal seed + two Newton steps (<= 1 ulp), no slow-path branch
        const double den = 2.0 + em;
        double inv;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(den));
        inv = fma(inv, fma(-den, inv, 1.0), inv);
        inv = fma(inv, fma(-den, inv, 1.0), inv);
#else
        const double inv = 1.0 / (2.0 + em);
#endif
        const double th = copysign(-em * inv, Y);
        const double sech2 = (4.0 * (1.0 + em)) * (inv * inv);
        double s1, c1, s2, c2, s3, c3;
#if B200CS_BICKLEY_WIDE && B200CS_BICKLEY_ONEGUARD
        {   // ONE range test for the three arguments instead of a guarded branch in each sincos: a step
            // attempt becomes a few large basic blocks (what gave the double gyre +4 %)
            const double a1 = p[6] * (y[0] - p[9] * tt), a2 = p[7] * (y[0] - p[10] * tt), a3 = p[8] * (y[0] - p[11] * tt);
            if (trig_in_range(a1) && trig_in_range(a2) && trig_in_range(a3)) {
                sincos_wide_core(a1, &s1, &c1);
                sincos_wide_core(a2, &s2, &c2);
                sincos_wide_core(a3, &s3, &c3);
            } else {
                const double2 q1 = sincos_slow(a1), q2 = sincos_slow(a2), q3 = sincos_slow(a3);
                s1 = q1.x; c1 = q1.y; s2 = q2.x; c2 = q2.y; s3 = q3.x; c3 = q3.y;
            }
        }
#elif B200CS_BICKLEY_WIDE
        sincos_wide(p[6] * (y[0] - p[9] * tt), &s1, &c1);
        sincos_wide(p[7] * (y[0] - p[10] * tt), &s2, &c2);
        sincos_wide(p[8] * (y[0] - p[11] * tt), &s3, &c3);
#else
        sincos_fast(p[6] * (y[0] - p[9] * tt), &s1, &c1);
        sincos_fast(p[7] * (y[0] - p[10] * tt), &s2, &c2);
        sincos_fast(p[8] * (y[0] - p[11] * tt), &s3, &c3);
#endif
        const double csum = fma(p[5], c3, fma(p[4], c2, p[3] * c1));
        const double ssum = fma(p[5] * p[8], s3, fma(p[4] * p[7], s2, p[3] * p[6] * s1));
        dy[0] = p[0] * fma(2.0 * p[1] * th * sech2, csum, p[1] * sech2);
        dy[1] = -p[0] * (p[1] * p[2] * sech2 * ssum);
    }
};

struct Abc {
    static constexpr int N = 3;
    static constexpr int kAux = B200CS_STRICT_RHS ? 0 : 1;
    const RhsParams &P;
    __device__ __forceinline__ explicit Abc(const RhsParams &P_) : P(P_) {}
    // A(t) = p1 + p4 * tt * sin(pi*tt)   (flows.py:1256-1257)
    template <int M>
    __device__ __forceinline__ void time_part(const double (&t)[M], double (&aux)[M]) const {
        const double *p = P.p;
        double arg[M], st[M];
#if B200CS_ABC_WIDE
#pragma unroll
        for (int m = 0; m < M; ++m) arg[m] = p[0] * t[m];
        sinpi12_v<M>(arg, st);   // sin(pi tt) with the exact reduction (no rounding of pi*tt)
#else
#pragma unroll
        for (int m = 0; m < M; ++m) arg[m] = kPi * (p[0] * t[m]);
        sin_v<M>(arg, st);
#endif
#pragma unroll
        for (int m = 0; m < M; ++m) aux[m] = fma(p[4] * (p[0] * t[m]), st[m], p[1]);
    }
    __device__ __forceinline__ void eval(double At, double t, const double (&y)[3], double (&dy)[3]) const {
        const double *p = P.p;
#if B200CS_STRICT_RHS
        {   // flows.py:1255-1258 token for token
            const double pi = kPi;
            const double tt = p[0] * t;
            dy[0] = p[0] * ((p[1] + p[4] * tt * sin(pi * tt)) * sin(y[2]) + p[3] * cos(y[1]));
            dy[1] = p[0] * (p[2] * sin(y[0]) + (p[1] + p[4] * tt * sin(pi * tt)) * cos(y[1]));
            dy[2] = p[0] * (p[3] * sin(y[1]) + p[2] * cos(y[1]));
            return;
        }
#endif
        const double arg[2] = {y[0], y[2]};
        double s02[2], s1, c1;
#if B200CS_ABC_WIDE
        sin_wide_v<2>(arg, s02);
        sincos_wide(y[1], &s1, &c1);
#else
        sin_v<2>(arg, s02);
        sincos_fast(y[1], &s1, &c1);
#endif
        dy[0] = p[0] * fma(At, s02[1], p[3] * c1);
        dy[1] = p[0] * fma(p[2], s02[0], At * c1);
        // the reference uses y[1] in both terms of dz (flows.py:1258); kept for parity
        dy[2] = p[0] * fma(p[3], s1, p[2] * c1);
    }
};

// Python float modulo: the result takes the sign of the (positive) divisor
__device__ __forceinline__ double pymod_pos(double a, double m) {
    double r = fmod(a, m);
    if (r < 0.0) r += m;
    return r;
}

// a % 360 (Python semantics) without fmod's loop and branches: k = rint(a / 360) by the magic
// constant, r = a - 360 k EXACTLY (one FMA: the exact remainder of a by 360 is representable), in
// [-180, 180], and a negative r moves up by 360 -- the same final addition (and the same rounding of
// it, e.g. -1e-20 % 360 == 360.0) as Python's fmod-then-adjust.  |a| >= 2^40 or non-finite: fmod.
__device__ __forceinline__ double pymod360(double a) {
    if (!(fabs(a) < 1.0e12)) return pymod_pos(a, 360.0);
    const double t = fma(a, 1.0 / 360.0, 6755399441055744.0);
    const double k = t - 6755399441055744.0;
    double r = fma(-360.0, k, a);
    // a / 360 within rounding of a half-integer can leave |r| a hair above 180: still exact, still in (-360, 360)
    if (r < 0.0) r += 360.0;
    return r;
}

// cos(x) with the wide kernel (one reduction by pi, even polynomial on |r| <= pi/2), no libm guard:
// the caller passes a latitude in radians
__device__ __forceinline__ double cos_wide_core(double x) {
    const double t = fma(x, kWide.inv_pi, kWide.magic);
    const int q = __double2loint(t);
    const double k = t - kWide.magic;
    const double r = fma(-k, kWide.pi_lo, fma(-k, kWide.pi_hi, x));
    const double z = r * r;
    double pc = kWide.cc[7];
#pragma unroll
    for (int j = 6; j >= 0; --j) pc = fma(pc, z, kWide.cc[j]);
    return flip_sign(fma(z, pc, 1.0), q & 1);
}

// LINEAR = true: get_flow_linear_2D (flows.py:418-506), trilinear eval_linear on the raw (u, v)
// data instead of the tri-cubic spline; everything else (longitude wrap, spherical scaling) is
// the same expression in the reference.
template <int SPHERICAL, bool LINEAR = false>
struct Spline2D {
    static constexpr int N = 2;
    static constexpr int kAux = 0;
    static constexpr bool kOutOfLine = B200CS_SPLINE_NOINLINE != 0;
    const RhsParams &P;
    __device__ __forceinline__ explicit Spline2D(const RhsParams &P_) : P(P_) {}
    template <int M>
    __device__ __forceinline__ void time_part(const double (&)[M], double (&)[M]) const {}
#if B200CS_SPLINE_NOINLINE
    // The 64-tap RHS is ~650 instructions; inlined into 12 stages + FSAL + hinit the attempt loop is
    // several times the 32 KB instruction cache, which is why this kernel ran in lockstep blocks.
    // Out of line there is one copy: the loop fits the cache and blocks can run free.
    static __device__ __noinline__ double2 eval_ool(const RhsParams *Pp, double t, double y0, double y1) {
        const Spline2D self(*Pp);
        const double y[2] = {y0, y1};
        double dy[2];
        self.eval_body(t, y, dy);
        return make_double2(dy[0], dy[1]);
    }
    static __device__ __noinline__ double2 eval_ool_slot(int slot, double t, double y0, double y1) {
        const Spline2D self(c_rhs_slots[slot]);
        const double y[2] = {y0, y1};
        double dy[2];
        self.eval_body(t, y, dy);
        return make_double2(dy[0], dy[1]);
    }
    __device__ __forceinline__ void eval(double, double t, const double (&y)[2], double (&dy)[2]) const {
        const double2 r = (P.slot >= 0) ? eval_ool_slot(P.slot, t, y[0], y[1]) : eval_ool(&P, t, y[0], y[1]);
        dy[0] = r.x;
        dy[1] = r.y;
    }
#else
    __device__ __forceinline__ void eval(double, double t, const double (&y)[2], double (&dy)[2]) const {
        eval_body(t, y, dy);
    }
#endif
    __device__ __forceinline__ void eval_body(double t, const double (&y)[2], double (&dy)[2]) const {
        const double p0 = P.p[0];
        double xx = y[0];
        const double yy = y[1];
#if B200CS_SPLINE_LEAN_SPH && !B200CS_STRICT_RHS
        if (SPHERICAL == 1) xx = pymod360(y[0] - 180.0) - 180.0;
        if (SPHERICAL == 2) xx = pymod360(y[0]);
#else
        if (SPHERICAL == 1) xx = pymod_pos(y[0] - 180.0, 360.0) - 180.0;
        if (SPHERICAL == 2) xx = pymod_pos(y[0], 360.0);
#endif
        double u, v;
        const double tt = p0 * t;
        // evaluations outside the data grid are counted inside the evaluators (spline.cuh, count_outside):
        // the only guard on the extrapolation modes, which no reference test pins
        if (LINEAR) eval_linear_uv(P.grid, P.coef_uv, tt, xx, yy, u, v);
        else eval_spline_uv(P.grid, P.coef_uv, tt, xx, yy, u, v);
        if (SPHERICAL) {
            // ((p0*u)*180) / (pi*r*cos(yy*pi/180))   (flows.py:165-196)
#if B200CS_STRICT_RHS
            dy[0] = ((p0 * u) * 180.0) / (kPi * P.r * cos(yy * kPi / 180.0));
            dy[1] = ((p0 * v) * 180.0) / (kPi * P.r);
#elif B200CS_SPLINE_LEAN_SPH
            // the same expressions with the divisions as reciprocal + one FMA residual correction (the
            // quotient to <= 1 ulp, no slow-path branch) and the cosine from the wide kernel; the
            // latitude argument yy*pi/180 keeps the reference's two roundings
            const double ypi = yy * kPi;
            const double ql = ypi * (1.0 / 180.0);
            const double lat = fma(fma(-ql, 180.0, ypi), 1.0 / 180.0, ql);   // (yy*pi)/180, correctly rounded
            const double cl = (fabs(lat) < 1.0e5) ? cos_wide_core(lat) : cos(lat);
            const double den = kPi * P.r * cl;
            double rd;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rd) : "d"(den));
            rd = fma(rd, fma(-den, rd, 1.0), rd);
            rd = fma(rd, fma(-den, rd, 1.0), rd);
            const double nu = (p0 * u) * 180.0;
            const double qu = nu * rd;
            dy[0] = fma(fma(-qu, den, nu), rd, qu);
            const double nv = (p0 * v) * 180.0;
            const double qv = nv * P.d[6];                 // 1 / (pi r), from the host
            dy[1] = fma(fma(-qv, P.d[7], nv), P.d[6], qv);  // P.d[7] = pi r
#else
            dy[0] = ((p0 * u) * 180.0) / (kPi * P.r * cos_fast(yy * kPi / 180.0));
            dy[1] = ((p0 * v) * 180.0) / (kPi * P.r);
#endif
        } else {
            dy[0] = p0 * u;
            dy[1] = p0 * v;
        }
    }
};

}  // namespace b200cs
