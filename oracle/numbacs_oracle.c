/*
 * numbacs_oracle.c -- CPU restatement of the NumbaCS flow-map + FTLE + LAVD hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (numbacs_b200/)
 * never links, imports or calls anything in this directory.
 *
 * What it restates (all paths relative to /root/reference):
 *   - RHS of the predefined flows            src/numbacs/flows.py:1146-1158 (double_gyre),
 *                                            1182-1213 (bickley_jet), 1249-1258 (abc)
 *   - cubic-spline RHS incl. spherical modes src/numbacs/flows.py:156-253
 *   - particle loops                         src/numbacs/integration.py:7-61, 64-120, 123-182, 467-533
 *   - FTLE                                   src/numbacs/diagnostics.py:21-65, utils.py:9-46, 168-189
 *   - LAVD + composite Simpson               src/numbacs/diagnostics.py:272-379, utils.py:611-655
 *   - aux-grid flow map                      src/numbacs/integration.py:249-464
 *   - Cauchy-Green tensor / eigen-pairs      src/numbacs/diagnostics.py:68-269, utils.py:49-124
 *     (np.linalg.eigh on 2x2 = LAPACK dlaev2, restated; pinned by Cevals/Cevecs*.npy and by the
 *     live numbacs.diagnostics.C_eig_2D, which imports here)
 *   - FTLE ridge points                      src/numbacs/extraction/ridges.py:9-76, 232-318
 *   - flow-map composition                   src/numbacs/integration.py:609-737 (bilinear
 *     interpolation.splines.eval_linear, CONSTANT extrapolation; pinned by fm_ci/fm_cs.npy)
 *
 * The arithmetic of the ODE solver and of the spline is NOT in the reference tree: it lives in
 * the third-party packages `numbalsoda` (unpinned, pyproject.toml:32) and `interpolation>=2.2.6`
 * (pyproject.toml:31), neither of which is installed here.  Their published algorithms are
 * restated instead:
 *   - numbalsoda.dop853 = Hairer/Noersett/Wanner DOP853 (dop853.f; C version by J. Colinge):
 *     classical controller (safe .9, fac1 .333, fac2 6, beta 0), hinit with iord 8, one
 *     continuous integration over t_eval with contd8 dense output at interior output times.
 *   - interpolation.splines.prefilter(k=3): separable natural-BC cubic B-spline prefilter;
 *     eval_spline(k=3): 4x4x4 tensor-product uniform cubic B-spline.
 * Parity is PINNED by the reference's own golden vectors (tests/testing_data/fm.npy, fm_n.npy,
 * ftle.npy, lavd.npy; spline tables of tests/test_flows.py) -- see tests/test_oracle_golden.py.
 * Parity is UNPINNED for: spline extrapolation outside the data grid, spherical=1/2 flow maps,
 * bickley/abc flow maps (no golden exists in the reference), nmax/stiffness exits.
 *
 * Floating point: compiled with -ffp-contract=off so every a*b+c is two roundings, as in the
 * reference's numba/LLVM (no fast-math) and numbalsoda builds; libm is glibc's.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "dop853_coefs.h"

#define MAXN 3

enum { FLOW_DOUBLE_GYRE = 0, FLOW_BICKLEY_JET = 1, FLOW_ABC = 2, FLOW_SPLINE2D = 3, FLOW_LINEAR2D = 4,
       FLOW_CALLBACK = 5 /* the RHS is a C function pointer (oracle/shims/numbalsoda) */ };
enum { EXTRAP_CONSTANT = 0, EXTRAP_LINEAR = 1, EXTRAP_NEAREST = 2 };

/* ------------------------------------------------------------------ spline ------------- */

typedef struct {
    double a[3], b[3];
    int64_t n[3];          /* data points per axis; coefficient array is (n+2) per axis */
    const double *C;       /* (n0+2, n1+2, n2+2) C-order */
    int extrap;
} spline3_t;

/* uniform cubic B-spline blending weights, Horner form of the matrix
 *   [-1 3 -3 1; 3 -6 0 4; -3 3 3 1; 1 0 0 0]/6 acting on [l^3 l^2 l 1];
 * outside [0,1] ('linear' extrapolation) the weights are continued linearly. */
static void bspline_weights(double l, int linear_ext, double *P)
{
    if (linear_ext && l < 0.0) {
        P[0] = (-3.0 / 6.0) * l + 1.0 / 6.0;
        P[1] = (0.0 / 6.0) * l + 4.0 / 6.0;
        P[2] = (3.0 / 6.0) * l + 1.0 / 6.0;
        P[3] = 0.0;
    } else if (linear_ext && l > 1.0) {
        double m = l - 1.0;
        P[0] = (3 * (-1.0 / 6.0) + 2 * (3.0 / 6.0) + (-3.0 / 6.0)) * m + (-1.0 / 6.0 + 3.0 / 6.0 - 3.0 / 6.0 + 1.0 / 6.0);
        P[1] = (3 * (3.0 / 6.0) + 2 * (-6.0 / 6.0) + 0.0) * m + (3.0 / 6.0 - 6.0 / 6.0 + 0.0 + 4.0 / 6.0);
        P[2] = (3 * (-3.0 / 6.0) + 2 * (3.0 / 6.0) + (3.0 / 6.0)) * m + (-3.0 / 6.0 + 3.0 / 6.0 + 3.0 / 6.0 + 1.0 / 6.0);
        P[3] = (3 * (1.0 / 6.0)) * m + (1.0 / 6.0);
    } else {
        double l2 = l * l, l3 = l2 * l;
        P[0] = (-1.0 / 6.0) * l3 + (3.0 / 6.0) * l2 + (-3.0 / 6.0) * l + 1.0 / 6.0;
        P[1] = (3.0 / 6.0) * l3 + (-6.0 / 6.0) * l2 + (0.0 / 6.0) * l + 4.0 / 6.0;
        P[2] = (-3.0 / 6.0) * l3 + (3.0 / 6.0) * l2 + (3.0 / 6.0) * l + 1.0 / 6.0;
        P[3] = (1.0 / 6.0) * l3;
    }
}

/* cell index + local coordinate on a uniform axis (a, b, n):
 *   delta=(b-a)/(n-1); i=clamp(floor((x-a)/delta), 0, n-2); lambda=((x-a)-i*delta)/delta */
static inline void axis_locate(double a, double b, int64_t n, double x, int64_t *i, double *lam)
{
    double delta = (b - a) / (double)(n - 1);
    double d = x - a;
    double fi = floor(d / delta);
    int64_t ii = (fi < 0.0) ? 0 : ((fi > (double)(n - 2)) ? (n - 2) : (int64_t)fi);
    *i = ii;
    *lam = (d - (double)ii * delta) / delta;
}

/* eval_spline(grid, C, [p0,p1,p2], k=3, extrap_mode): 64-tap tensor product. */
double oracle_eval_spline3(const spline3_t *s, double p0, double p1, double p2)
{
    double p[3] = {p0, p1, p2};
    int64_t idx[3];
    double P[3][4];
    for (int d = 0; d < 3; ++d) {
        double x = p[d];
        if (s->extrap == EXTRAP_CONSTANT) {
            if (x < s->a[d] || x > s->b[d]) return 0.0;
        } else if (s->extrap == EXTRAP_NEAREST) {
            x = fmax(s->a[d], fmin(s->b[d], x));
        }
        double lam;
        axis_locate(s->a[d], s->b[d], s->n[d], x, &idx[d], &lam);
        bspline_weights(lam, s->extrap == EXTRAP_LINEAR, P[d]);
    }
    int64_t s1 = s->n[2] + 2, s0 = (s->n[1] + 2) * s1;
    const double *base = s->C + idx[0] * s0 + idx[1] * s1 + idx[2];
    double acc0 = 0.0;
    for (int a = 0; a < 4; ++a) {
        double acc1 = 0.0;
        for (int b = 0; b < 4; ++b) {
            const double *c = base + a * s0 + b * s1;
            double acc2 = P[2][0] * c[0] + P[2][1] * c[1] + P[2][2] * c[2] + P[2][3] * c[3];
            acc1 += P[1][b] * acc2;
        }
        acc0 += P[0][a] * acc1;
    }
    return acc0;
}

/* eval_linear(grid, F, [p0,p1,p2], extrap): trilinear on the (n0,n1,n2) data array. */
double oracle_eval_linear3(const spline3_t *s, double p0, double p1, double p2)
{
    double p[3] = {p0, p1, p2};
    int64_t idx[3];
    double lam[3];
    for (int d = 0; d < 3; ++d) {
        double x = p[d];
        if (s->extrap == EXTRAP_CONSTANT) {
            if (x < s->a[d] || x > s->b[d]) return 0.0;
        } else if (s->extrap == EXTRAP_NEAREST) {
            x = fmax(s->a[d], fmin(s->b[d], x));
        }
        axis_locate(s->a[d], s->b[d], s->n[d], x, &idx[d], &lam[d]);
    }
    int64_t s1 = s->n[2], s0 = s->n[1] * s1;
    const double *c = s->C + idx[0] * s0 + idx[1] * s1 + idx[2];
    double v = 0.0;
    for (int a = 0; a < 2; ++a) {
        double wa = a ? lam[0] : 1.0 - lam[0];
        double va = 0.0;
        for (int b = 0; b < 2; ++b) {
            double wb = b ? lam[1] : 1.0 - lam[1];
            const double *cc = c + a * s0 + b * s1;
            va += wb * ((1.0 - lam[2]) * cc[0] + lam[2] * cc[1]);
        }
        v += wa * va;
    }
    return v;
}

/* 1-D natural-BC cubic B-spline prefilter: data d[0..n-1] (stride sd) -> coefs c[0..n+1] (stride sc).
 *   c[0]-2c[1]+c[2]=0 ; (c[i]+4c[i+1]+c[i+2])/6=d[i] ; c[n-1]-2c[n]+c[n+1]=0.
 * The natural conditions give c[1]=d[0], c[n]=d[n-1]; the interior is a (1,4,1) Thomas solve. */
static void prefilter_1d(const double *d, int64_t sd, int64_t n, double *c, int64_t sc, double *work)
{
    /* unknowns u[k]=c[k+1], k=0..n-1 ; u[0]=d[0], u[n-1]=d[n-1] */
    double *cp = work, *dp = work + n;
    c[1 * sc] = d[0];
    c[n * sc] = d[(n - 1) * sd];
    if (n > 2) {
        /* rows k=1..n-2:  u[k-1] + 4u[k] + u[k+1] = 6 d[k] */
        int64_t m = n - 2;
        for (int64_t k = 0; k < m; ++k) {
            double rhs = 6.0 * d[(k + 1) * sd];
            if (k == 0) rhs -= d[0];
            if (k == m - 1) rhs -= d[(n - 1) * sd];
            double lower = (k == 0) ? 0.0 : 1.0;
            double denom = 4.0 - lower * (k ? cp[k - 1] : 0.0);
            cp[k] = 1.0 / denom;
            dp[k] = (rhs - lower * (k ? dp[k - 1] : 0.0)) / denom;
        }
        /* unknown k is u[k+1] = c[k+2] */
        c[(m + 1) * sc] = dp[m - 1];
        for (int64_t k = m - 2; k >= 0; --k) c[(k + 2) * sc] = dp[k] - cp[k] * c[(k + 3) * sc];
    }
    c[0] = 2.0 * c[1 * sc] - c[2 * sc];
    c[(n + 1) * sc] = 2.0 * c[n * sc] - c[(n - 1) * sc];
}

/* prefilter(grid, data[n0,n1,n2], k=3) -> coefs[(n0+2),(n1+2),(n2+2)] (flows.py:43-44, 116) */
void oracle_prefilter3(const double *data, int64_t n0, int64_t n1, int64_t n2, double *coefs)
{
    int64_t m0 = n0 + 2, m1 = n1 + 2, m2 = n2 + 2;
    int64_t nmax = n0 > n1 ? n0 : n1;
    if (n2 > nmax) nmax = n2;
    memset(coefs, 0, sizeof(double) * m0 * m1 * m2);
    /* axis 2: data rows -> coefs[i0+1, i1+1, :] */
#pragma omp parallel
    {
        double *work = (double *)malloc(sizeof(double) * 3 * (nmax + 2));
        double *tmp = work + 2 * (nmax + 2);
#pragma omp for collapse(2)
        for (int64_t i0 = 0; i0 < n0; ++i0)
            for (int64_t i1 = 0; i1 < n1; ++i1)
                prefilter_1d(data + (i0 * n1 + i1) * n2, 1, n2,
                             coefs + ((i0 + 1) * m1 + (i1 + 1)) * m2, 1, work);
        /* axis 1: in place on coefs[i0+1, 1..n1, i2] -> coefs[i0+1, 0..n1+1, i2] */
#pragma omp for collapse(2)
        for (int64_t i0 = 0; i0 < n0; ++i0)
            for (int64_t i2 = 0; i2 < m2; ++i2) {
                double *col = coefs + (i0 + 1) * m1 * m2 + i2;
                for (int64_t k = 0; k < n1; ++k) tmp[k] = col[(k + 1) * m2];
                prefilter_1d(tmp, 1, n1, col, m2, work);
            }
        /* axis 0 */
#pragma omp for collapse(2)
        for (int64_t i1 = 0; i1 < m1; ++i1)
            for (int64_t i2 = 0; i2 < m2; ++i2) {
                double *col = coefs + i1 * m2 + i2;
                for (int64_t k = 0; k < n0; ++k) tmp[k] = col[(k + 1) * m1 * m2];
                prefilter_1d(tmp, 1, n0, col, m1 * m2, work);
            }
        free(work);
    }
}

/* ------------------------------------------------------------------ flows -------------- */

typedef struct {
    int kind;
    int ndim;
    int spherical;   /* 0,1,2  (flows.py:137-143) */
    double r;
    spline3_t u, v;
    void (*callback)(double, double *, double *, double *);   /* FLOW_CALLBACK: lsoda_sig */
} flow_t;

/* Python/numba float modulo: result has the sign of the divisor (flows.py:162, 205) */
static inline double pymod(double a, double m)
{
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}

static void rhs_eval(const flow_t *f, double t, const double *y, double *dy, const double *p)
{
    const double pi = 3.141592653589793;
    switch (f->kind) {
    case FLOW_CALLBACK:
        f->callback(t, (double *)y, dy, (double *)p);
        break;
    case FLOW_DOUBLE_GYRE: { /* flows.py:1152-1158 */
        double tt = p[0] * t;
        double a = p[2] * sin(p[4] * tt + p[5]);
        double b = 1 - 2 * a;
        double ff = a * (y[0] * y[0]) + b * y[0];
        double df = 2 * a * y[0] + b;
        dy[0] = p[0] * (-pi * p[1] * sin(pi * ff) * cos(pi * y[1]) - p[3] * y[0]);
        dy[1] = p[0] * (pi * p[1] * cos(pi * ff) * sin(pi * y[1]) * df - p[3] * y[1]);
        break;
    }
    case FLOW_BICKLEY_JET: { /* flows.py:1189-1213 */
        double tt = p[0] * t;
        double Y = y[1] / p[2];
        double ch = cosh(Y);
        double sech2 = 1 / (ch * ch);
        dy[0] = p[0] * (p[1] * sech2 +
                        2 * p[1] * tanh(Y) * sech2 *
                            (p[3] * cos(p[6] * (y[0] - p[9] * tt)) +
                             p[4] * cos(p[7] * (y[0] - p[10] * tt)) +
                             p[5] * cos(p[8] * (y[0] - p[11] * tt))));
        dy[1] = -p[0] * (p[1] * p[2] * sech2 *
                         (p[3] * p[6] * sin(p[6] * (y[0] - p[9] * tt)) +
                          p[4] * p[7] * sin(p[7] * (y[0] - p[10] * tt)) +
                          p[5] * p[8] * sin(p[8] * (y[0] - p[11] * tt))));
        break;
    }
    case FLOW_ABC: { /* flows.py:1255-1258 (dz uses y[1] in both terms, as the reference does) */
        double tt = p[0] * t;
        double At = p[1] + p[4] * tt * sin(pi * tt);
        dy[0] = p[0] * (At * sin(y[2]) + p[3] * cos(y[1]));
        dy[1] = p[0] * (p[2] * sin(y[0]) + At * cos(y[1]));
        dy[2] = p[0] * (p[3] * sin(y[1]) + p[2] * cos(y[1]));
        break;
    }
    case FLOW_SPLINE2D:   /* flows.py:156-253 */
    case FLOW_LINEAR2D: { /* flows.py:458-503: same expressions with eval_linear on the raw data */
        double tt = p[0] * t;
        double xx = y[0], yy = y[1];
        if (f->spherical == 1) xx = pymod(y[0] - 180, 360) - 180;
        else if (f->spherical == 2) xx = pymod(y[0], 360);
        double u, v;
        if (f->kind == FLOW_LINEAR2D) {
            u = oracle_eval_linear3(&f->u, tt, xx, yy);
            v = oracle_eval_linear3(&f->v, tt, xx, yy);
        } else {
            u = oracle_eval_spline3(&f->u, tt, xx, yy);
            v = oracle_eval_spline3(&f->v, tt, xx, yy);
        }
        if (f->spherical) {
            dy[0] = ((p[0] * u) * 180) / (pi * f->r * cos(yy * pi / 180));
            dy[1] = ((p[0] * v) * 180) / (pi * f->r);
        } else {
            dy[0] = p[0] * u;
            dy[1] = p[0] * v;
        }
        break;
    }
    }
}

void oracle_rhs(const flow_t *f, double t, const double *y, double *dy, const double *p)
{
    rhs_eval(f, t, y, dy, p);
}

/* ------------------------------------------------------------------ DOP853 ------------- */

static inline double sgn(double a, double b) { return (b < 0.0) ? -fabs(a) : fabs(a); }

typedef struct {
    int64_t nfev, naccept, nreject, nstep;
} dopstat_t;

static double hinit(const flow_t *f, int n, double x, const double *y, double posneg,
                    const double *f0, double hmax, double rtol, double atol, const double *p)
{
    double dnf = 0.0, dny = 0.0, f1[MAXN], yy1[MAXN];
    for (int i = 0; i < n; ++i) {
        double sk = atol + rtol * fabs(y[i]);
        double sqr = f0[i] / sk;
        dnf += sqr * sqr;
        sqr = y[i] / sk;
        dny += sqr * sqr;
    }
    double h;
    if (dnf <= 1.0e-10 || dny <= 1.0e-10) h = 1.0e-6;
    else h = sqrt(dny / dnf) * 0.01;
    h = fmin(h, hmax);
    h = sgn(h, posneg);
    for (int i = 0; i < n; ++i) yy1[i] = y[i] + h * f0[i];
    rhs_eval(f, x + h, yy1, f1, p);
    double der2 = 0.0;
    for (int i = 0; i < n; ++i) {
        double sk = atol + rtol * fabs(y[i]);
        double sqr = (f1[i] - f0[i]) / sk;
        der2 += sqr * sqr;
    }
    der2 = sqrt(der2) / h;
    double der12 = fmax(fabs(der2), sqrt(dnf));
    double h1;
    if (der12 <= 1.0e-15) h1 = fmax(1.0e-6, fabs(h) * 1.0e-3);
    else h1 = pow(0.01 / der12, 1.0 / 8.0);
    h = fmin(100.0 * fabs(h), fmin(h1, hmax));
    return sgn(h, posneg);
}

/* numbalsoda.dop853(funcptr, u0, t_eval, rtol, atol, data) -> (usol[nt, n], success)
 * (call sites: integration.py:49, 108, 169, 520).  Returns 1 on success, <0 on failure
 * (-2 nmax exceeded, -3 step size underflow); `usol` rows not reached are left untouched. */
int oracle_dop853(const flow_t *f, int n, const double *u0, const double *t_eval, int64_t nt,
                  double rtol, double atol, const double *p, double *usol, dopstat_t *st)
{
    const double uround = 2.3e-16, safe = 0.9, fac1 = 0.333, fac2 = 6.0, beta = 0.0;
    const int64_t nmax = 100000;
    double y[MAXN], k1[MAXN], k2[MAXN], k3[MAXN], k4[MAXN], k5[MAXN], k6[MAXN], k7[MAXN],
        k8[MAXN], k9[MAXN], k10[MAXN], yy1[MAXN];
    double rc[8][MAXN];
    double x = t_eval[0], xend = t_eval[nt - 1];
    double facold = 1.0e-4, expo1 = 1.0 / 8.0 - beta * 0.2, facc1 = 1.0 / fac1, facc2 = 1.0 / fac2;
    double posneg = sgn(1.0, xend - x);
    double hmax = fabs(xend - x);
    int last = 0, reject = 0;
    int64_t iout = 1; /* next row of t_eval to emit */
    dopstat_t s = {0, 0, 0, 0};

    for (int i = 0; i < n; ++i) { y[i] = u0[i]; usol[i] = u0[i]; }
    if (nt < 2 || xend == x) { if (st) *st = s; return 1; }
    rhs_eval(f, x, y, k1, p);
    double h = hinit(f, n, x, y, posneg, k1, hmax, rtol, atol, p);
    s.nfev += 2;
    const int dense = (nt > 2);

    for (;;) {
        if (s.nstep > nmax) { if (st) *st = s; return -2; }
        if (0.1 * fabs(h) <= fabs(x) * uround) { if (st) *st = s; return -3; }
        if ((x + 1.01 * h - xend) * posneg > 0.0) { h = xend - x; last = 1; }
        s.nstep++;
        /* the twelve stages */
        for (int i = 0; i < n; ++i) yy1[i] = y[i] + h * DOP_A2_1 * k1[i];
        rhs_eval(f, x + DOP_C2 * h, yy1, k2, p);
        for (int i = 0; i < n; ++i) yy1[i] = y[i] + h * (DOP_A3_1 * k1[i] + DOP_A3_2 * k2[i]);
        rhs_eval(f, x + DOP_C3 * h, yy1, k3, p);
        for (int i = 0; i < n; ++i) yy1[i] = y[i] + h * (DOP_A4_1 * k1[i] + DOP_A4_3 * k3[i]);
        rhs_eval(f, x + DOP_C4 * h, yy1, k4, p);
        for (int i = 0; i < n; ++i)
            yy1[i] = y[i] + h * (DOP_A5_1 * k1[i] + DOP_A5_3 * k3[i] + DOP_A5_4 * k4[i]);
        rhs_eval(f, x + DOP_C5 * h, yy1, k5, p);
        for (int i = 0; i < n; ++i)
            yy1[i] = y[i] + h * (DOP_A6_1 * k1[i] + DOP_A6_4 * k4[i] + DOP_A6_5 * k5[i]);
        rhs_eval(f, x + DOP_C6 * h, yy1, k6, p);
        for (int i = 0; i < n; ++i)
            yy1[i] = y[i] + h * (DOP_A7_1 * k1[i] + DOP_A7_4 * k4[i] + DOP_A7_5 * k5[i] + DOP_A7_6 * k6[i]);
        rhs_eval(f, x + DOP_C7 * h, yy1, k7, p);
        for (int i = 0; i < n; ++i)
            yy1[i] = y[i] + h * (DOP_A8_1 * k1[i] + DOP_A8_4 * k4[i] + DOP_A8_5 * k5[i] + DOP_A8_6 * k6[i] +
                                 DOP_A8_7 * k7[i]);
        rhs_eval(f, x + DOP_C8 * h, yy1, k8, p);
        for (int i = 0; i < n; ++i)
            yy1[i] = y[i] + h * (DOP_A9_1 * k1[i] + DOP_A9_4 * k4[i] + DOP_A9_5 * k5[i] + DOP_A9_6 * k6[i] +
                                 DOP_A9_7 * k7[i] + DOP_A9_8 * k8[i]);
        rhs_eval(f, x + DOP_C9 * h, yy1, k9, p);
        for (int i = 0; i < n; ++i)
            yy1[i] = y[i] + h * (DOP_A10_1 * k1[i] + DOP_A10_4 * k4[i] + DOP_A10_5 * k5[i] + DOP_A10_6 * k6[i] +
                                 DOP_A10_7 * k7[i] + DOP_A10_8 * k8[i] + DOP_A10_9 * k9[i]);
        rhs_eval(f, x + DOP_C10 * h, yy1, k10, p);
        for (int i = 0; i < n; ++i)
            yy1[i] = y[i] + h * (DOP_A11_1 * k1[i] + DOP_A11_4 * k4[i] + DOP_A11_5 * k5[i] + DOP_A11_6 * k6[i] +
                                 DOP_A11_7 * k7[i] + DOP_A11_8 * k8[i] + DOP_A11_9 * k9[i] + DOP_A11_10 * k10[i]);
        rhs_eval(f, x + DOP_C11 * h, yy1, k2, p);
        double xph = x + h;
        for (int i = 0; i < n; ++i)
            yy1[i] = y[i] + h * (DOP_A12_1 * k1[i] + DOP_A12_4 * k4[i] + DOP_A12_5 * k5[i] + DOP_A12_6 * k6[i] +
                                 DOP_A12_7 * k7[i] + DOP_A12_8 * k8[i] + DOP_A12_9 * k9[i] + DOP_A12_10 * k10[i] +
                                 DOP_A12_11 * k2[i]);
        rhs_eval(f, xph, yy1, k3, p);
        s.nfev += 11;
        for (int i = 0; i < n; ++i) {
            k4[i] = DOP_B1 * k1[i] + DOP_B6 * k6[i] + DOP_B7 * k7[i] + DOP_B8 * k8[i] + DOP_B9 * k9[i] +
                    DOP_B10 * k10[i] + DOP_B11 * k2[i] + DOP_B12 * k3[i];
            k5[i] = y[i] + h * k4[i];
        }
        /* error estimation */
        double err = 0.0, err2 = 0.0;
        for (int i = 0; i < n; ++i) {
            double sk = atol + rtol * fmax(fabs(y[i]), fabs(k5[i]));
            double erri = k4[i] - DOP_BHH1 * k1[i] - DOP_BHH2 * k9[i] - DOP_BHH3 * k3[i];
            double sqr = erri / sk;
            err2 += sqr * sqr;
            erri = DOP_ER1 * k1[i] + DOP_ER6 * k6[i] + DOP_ER7 * k7[i] + DOP_ER8 * k8[i] + DOP_ER9 * k9[i] +
                   DOP_ER10 * k10[i] + DOP_ER11 * k2[i] + DOP_ER12 * k3[i];
            sqr = erri / sk;
            err += sqr * sqr;
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / (deno * (double)n));
        double fac11 = pow(err, expo1);
        double fac = fac11 / pow(facold, beta);
        fac = fmax(facc2, fmin(facc1, fac / safe));
        double hnew = h / fac;
        if (err <= 1.0) {
            facold = fmax(err, 1.0e-4);
            s.naccept++;
            rhs_eval(f, xph, k5, k4, p);
            s.nfev++;
            /* dense output only when an interior output time falls inside (x, xph] */
            int need_dense = dense && iout < nt - 1 && (t_eval[iout] - xph) * posneg <= 0.0;
            if (need_dense) {
                for (int i = 0; i < n; ++i) {
                    rc[0][i] = y[i];
                    double ydiff = k5[i] - y[i];
                    rc[1][i] = ydiff;
                    double bspl = h * k1[i] - ydiff;
                    rc[2][i] = bspl;
                    rc[3][i] = ydiff - h * k4[i] - bspl;
                    rc[4][i] = DOP_D4_1 * k1[i] + DOP_D4_6 * k6[i] + DOP_D4_7 * k7[i] + DOP_D4_8 * k8[i] +
                               DOP_D4_9 * k9[i] + DOP_D4_10 * k10[i] + DOP_D4_11 * k2[i] + DOP_D4_12 * k3[i];
                    rc[5][i] = DOP_D5_1 * k1[i] + DOP_D5_6 * k6[i] + DOP_D5_7 * k7[i] + DOP_D5_8 * k8[i] +
                               DOP_D5_9 * k9[i] + DOP_D5_10 * k10[i] + DOP_D5_11 * k2[i] + DOP_D5_12 * k3[i];
                    rc[6][i] = DOP_D6_1 * k1[i] + DOP_D6_6 * k6[i] + DOP_D6_7 * k7[i] + DOP_D6_8 * k8[i] +
                               DOP_D6_9 * k9[i] + DOP_D6_10 * k10[i] + DOP_D6_11 * k2[i] + DOP_D6_12 * k3[i];
                    rc[7][i] = DOP_D7_1 * k1[i] + DOP_D7_6 * k6[i] + DOP_D7_7 * k7[i] + DOP_D7_8 * k8[i] +
                               DOP_D7_9 * k9[i] + DOP_D7_10 * k10[i] + DOP_D7_11 * k2[i] + DOP_D7_12 * k3[i];
                }
                for (int i = 0; i < n; ++i)
                    yy1[i] = y[i] + h * (DOP_A14_1 * k1[i] + DOP_A14_7 * k7[i] + DOP_A14_8 * k8[i] +
                                         DOP_A14_9 * k9[i] + DOP_A14_10 * k10[i] + DOP_A14_11 * k2[i] +
                                         DOP_A14_12 * k3[i] + DOP_A14_13 * k4[i]);
                rhs_eval(f, x + DOP_C14 * h, yy1, k10, p);
                for (int i = 0; i < n; ++i)
                    yy1[i] = y[i] + h * (DOP_A15_1 * k1[i] + DOP_A15_6 * k6[i] + DOP_A15_7 * k7[i] +
                                         DOP_A15_8 * k8[i] + DOP_A15_11 * k2[i] + DOP_A15_12 * k3[i] +
                                         DOP_A15_13 * k4[i] + DOP_A15_14 * k10[i]);
                rhs_eval(f, x + DOP_C15 * h, yy1, k2, p);
                for (int i = 0; i < n; ++i)
                    yy1[i] = y[i] + h * (DOP_A16_1 * k1[i] + DOP_A16_6 * k6[i] + DOP_A16_7 * k7[i] +
                                         DOP_A16_8 * k8[i] + DOP_A16_9 * k9[i] + DOP_A16_13 * k4[i] +
                                         DOP_A16_14 * k10[i] + DOP_A16_15 * k2[i]);
                rhs_eval(f, x + DOP_C16 * h, yy1, k3, p);
                s.nfev += 3;
                for (int i = 0; i < n; ++i) {
                    rc[4][i] = h * (rc[4][i] + DOP_D4_13 * k4[i] + DOP_D4_14 * k10[i] + DOP_D4_15 * k2[i] + DOP_D4_16 * k3[i]);
                    rc[5][i] = h * (rc[5][i] + DOP_D5_13 * k4[i] + DOP_D5_14 * k10[i] + DOP_D5_15 * k2[i] + DOP_D5_16 * k3[i]);
                    rc[6][i] = h * (rc[6][i] + DOP_D6_13 * k4[i] + DOP_D6_14 * k10[i] + DOP_D6_15 * k2[i] + DOP_D6_16 * k3[i]);
                    rc[7][i] = h * (rc[7][i] + DOP_D7_13 * k4[i] + DOP_D7_14 * k10[i] + DOP_D7_15 * k2[i] + DOP_D7_16 * k3[i]);
                }
                while (iout < nt - 1 && (t_eval[iout] - xph) * posneg <= 0.0) {
                    double th = (t_eval[iout] - x) / h, th1 = 1.0 - th;
                    for (int i = 0; i < n; ++i)
                        usol[iout * n + i] =
                            rc[0][i] + th * (rc[1][i] + th1 * (rc[2][i] + th * (rc[3][i] + th1 * (rc[4][i] +
                                       th * (rc[5][i] + th1 * (rc[6][i] + th * rc[7][i]))))));
                    ++iout;
                }
            }
            for (int i = 0; i < n; ++i) { k1[i] = k4[i]; y[i] = k5[i]; }
            x = xph;
            if (last) {
                for (int i = 0; i < n; ++i) usol[(nt - 1) * n + i] = y[i];
                if (st) *st = s;
                return 1;
            }
            if (fabs(hnew) > hmax) hnew = posneg * hmax;
            if (reject) hnew = posneg * fmin(fabs(hnew), fabs(h));
            reject = 0;
        } else {
            hnew = h / fmin(facc1, fac11 / safe);
            reject = 1;
            if (s.naccept >= 1) s.nreject++;
            last = 0;
        }
        h = hnew;
    }
}

/* numbalsoda.dop853 with the RHS given as a C function pointer (the address of a numba @cfunc with
 * signature lsoda_sig): what oracle/shims/numbalsoda calls so that the UNMODIFIED reference source
 * runs on the same integrator.  Reentrant (the reference calls it from numba prange threads). */
int oracle_dop853_callback(void (*rhs)(double, double *, double *, double *), int n, const double *u0,
                           const double *t_eval, int64_t nt, double rtol, double atol, const double *p,
                           double *usol)
{
    flow_t f;
    memset(&f, 0, sizeof(f));
    f.kind = FLOW_CALLBACK;
    f.callback = rhs;
    if (n < 1 || n > MAXN) return -1;
    return oracle_dop853(&f, n, u0, t_eval, nt, rtol, atol, p, usol, NULL);
}

/* ------------------------------------------------------------------ particle loops ------ */

/* numpy/numba linspace as used at integration.py:44, 102, 164, 514: start + i*step, last = stop */
static void linspace_scaled(double p0, double t0, double t1, int64_t n, double *out)
{
    double step = (n > 1) ? (t1 - t0) / (double)(n - 1) : 0.0;
    for (int64_t i = 0; i < n; ++i) out[i] = p0 * (t0 + (double)i * step);
    if (n > 1) out[n - 1] = p0 * t1;
}

/* flowmap / flowmap_n (integration.py:7-61, 64-120): pts[npts, ndim] -> out[npts, n, ndim]
 * (n == 2 with last_only=1 reproduces flowmap: out[npts, ndim] = last row). */
int oracle_flowmap_pts(const flow_t *f, double t0, double T, const double *pts, int64_t npts,
                       const double *params, double rtol, double atol, const uint8_t *mask,
                       int64_t n, int last_only, double *out, double *tspan, int32_t *status,
                       int32_t *steps /* [npts,2] accept,reject */, int64_t *stats)
{
    int nd = f->ndim;
    double *t_eval = (double *)malloc(sizeof(double) * n);
    linspace_scaled(params[0], t0, t0 + T, n, t_eval);
    if (tspan) for (int64_t k = 0; k < n; ++k) tspan[k] = params[0] * t_eval[k]; /* integration.py:120 */
    int64_t tot_fev = 0, tot_acc = 0, tot_rej = 0;
    int64_t row = last_only ? nd : n * nd;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tot_fev, tot_acc, tot_rej)
    for (int64_t q = 0; q < npts; ++q) {
        double *o = out + q * row;
        for (int64_t k = 0; k < row; ++k) o[k] = 0.0;
        if (status) status[q] = 0;
        if (steps) { steps[2 * q] = 0; steps[2 * q + 1] = 0; }
        if (mask && mask[q]) continue;
        double usol_stack[64];
        double *usol = (n * nd <= 64) ? usol_stack : (double *)malloc(sizeof(double) * n * nd);
        for (int64_t k = 0; k < n * nd; ++k) usol[k] = 0.0;
        dopstat_t s;
        int rc = oracle_dop853(f, nd, pts + q * nd, t_eval, n, rtol, atol, params, usol, &s);
        if (last_only) for (int i = 0; i < nd; ++i) o[i] = usol[(n - 1) * nd + i];
        else for (int64_t k = 0; k < n * nd; ++k) o[k] = usol[k];
        if (usol != usol_stack) free(usol);
        if (status) status[q] = rc;
        if (steps) { steps[2 * q] = (int32_t)s.naccept; steps[2 * q + 1] = (int32_t)s.nreject; }
        tot_fev += s.nfev; tot_acc += s.naccept; tot_rej += s.nstep - s.naccept;
    }
    if (stats) { stats[0] = tot_fev; stats[1] = tot_acc; stats[2] = tot_rej; }
    free(t_eval);
    return 0;
}

/* flowmap_grid_2D / flowmap_n_grid_2D (integration.py:123-182, 467-533): 'ij' grid */
int oracle_flowmap_grid_2d(const flow_t *f, double t0, double T, const double *x, int64_t nx,
                           const double *y, int64_t ny, const double *params, double rtol,
                           double atol, const uint8_t *mask, int64_t n, int last_only, double *out,
                           double *tspan, int32_t *status, int32_t *steps, int64_t *stats)
{
    double *pts = (double *)malloc(sizeof(double) * 2 * nx * ny);
    for (int64_t i = 0; i < nx; ++i)
        for (int64_t j = 0; j < ny; ++j) {
            pts[2 * (i * ny + j)] = x[i];
            pts[2 * (i * ny + j) + 1] = y[j];
        }
    int rc = oracle_flowmap_pts(f, t0, T, pts, nx * ny, params, rtol, atol, mask, n, last_only,
                                out, tspan, status, steps, stats);
    free(pts);
    return rc;
}

/* ------------------------------------------------------------------ FTLE --------------- */

/* ftle_grid_2D (diagnostics.py:21-65) with gradF_stencil_2D (utils.py:41-44) and
 * eigvalsh_max_2D (utils.py:185-189) */
void oracle_ftle_grid_2d(const double *fm, int64_t nx, int64_t ny, double T, double dx, double dy,
                         const uint8_t *mask, double *ftle)
{
    double scaling = 1 / (2 * fabs(T));
    memset(ftle, 0, sizeof(double) * nx * ny);
#pragma omp parallel for schedule(static)
    for (int64_t i = 1; i < nx - 1; ++i)
        for (int64_t j = 1; j < ny - 1; ++j) {
            if (mask && mask[i * ny + j]) continue;
#define F(I, J, C) fm[(((I)*ny) + (J)) * 2 + (C)]
            double dxdx = (F(i + 1, j, 0) - F(i - 1, j, 0)) / (2 * dx);
            double dxdy = (F(i, j + 1, 0) - F(i, j - 1, 0)) / (2 * dy);
            double dydx = (F(i + 1, j, 1) - F(i - 1, j, 1)) / (2 * dx);
            double dydy = (F(i, j + 1, 1) - F(i, j - 1, 1)) / (2 * dy);
#undef F
            double off = dxdx * dxdy + dydx * dydy;
            double a = dxdx * dxdx + dydx * dydx, d = dxdy * dxdy + dydy * dydy;
            double trace = a + d;
            double disc = sqrt((a - d) * (a - d) + 4 * (off * off));
            double max_eig = 0.5 * (trace + disc);
            if (max_eig > 1) ftle[i * ny + j] = scaling * log(max_eig);
        }
}

/* ------------------------------------------------------------------ LAVD --------------- */

/* composite_simpsons (utils.py:611-655) */
double oracle_composite_simpsons(const double *f, int64_t len, double h)
{
    int64_t n = len - 1;
    double val;
    if (n % 2 == 0) {
        val = f[0];
        val += f[n];
        for (int64_t k = 1; k < n; ++k) val += (k % 2 != 0) ? 4 * f[k] : 2 * f[k];
        val *= h / 3;
    } else {
        n -= 1;
        val = f[0];
        val += f[len - 2];
        for (int64_t k = 1; k < n; ++k) val += (k % 2 != 0) ? 4 * f[k] : 2 * f[k];
        val *= h / 3;
        val += (5 * h / 12) * f[len - 1] + (2 * h / 3) * f[len - 2] - (h / 12) * f[len - 3];
    }
    return val;
}

/* vort_interp(pts) for an (npts, 3) array of (t, x, y) rows (flows.py:409, 631), all host threads */
void oracle_scalar_eval_many(const spline3_t *s, int linear, const double *pts, int64_t npts, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < npts; ++q)
        out[q] = linear ? oracle_eval_linear3(s, pts[3 * q], pts[3 * q + 1], pts[3 * q + 2])
                        : oracle_eval_spline3(s, pts[3 * q], pts[3 * q + 1], pts[3 * q + 2]);
}

/* lavd_grid_2D (diagnostics.py:272-379).  The vorticity interpolant is either the cubic spline
 * (get_callable_scalar, flows.py:387-415) or the trilinear one (get_callable_scalar_linear,
 * flows.py:601-636; used by the reference's own test, tests/test_diagnostics.py:153-173). */
void oracle_lavd_grid_2d(const double *fm_n, int64_t nx, int64_t ny, int64_t n, const double *tspan,
                         const spline3_t *vort, int linear, const double *xrav, const double *yrav,
                         double period_x, double period_y, const uint8_t *mask, double *lavd)
{
    int64_t npts = nx * ny;
    double *vavg = (double *)malloc(sizeof(double) * n);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        /* np.mean = pairwise summation in numpy, plain loop in numba; the difference is
         * O(1e-16) relative and far below every gate */
        double s = 0.0;
        for (int64_t q = 0; q < npts; ++q)
            s += linear ? oracle_eval_linear3(vort, tspan[k], xrav[q], yrav[q])
                        : oracle_eval_spline3(vort, tspan[k], xrav[q], yrav[q]);
        vavg[k] = s / (double)npts;
    }
    double dt = fabs(tspan[1] - tspan[0]);
    memset(lavd, 0, sizeof(double) * npts);
#pragma omp parallel
    {
        double *integrand = (double *)malloc(sizeof(double) * n);
#pragma omp for schedule(static)
        for (int64_t q = 0; q < npts; ++q) {
            if (mask && mask[q]) continue;
            for (int64_t k = 0; k < n; ++k) {
                double px = fm_n[(q * n + k) * 2], py = fm_n[(q * n + k) * 2 + 1];
                if (period_x != 0.0) px = pymod(px, period_x);
                if (period_y != 0.0) py = pymod(py, period_y);
                double w = linear ? oracle_eval_linear3(vort, tspan[k], px, py)
                                  : oracle_eval_spline3(vort, tspan[k], px, py);
                integrand[k] = fabs(w - vavg[k]);
            }
            lavd[q] = oracle_composite_simpsons(integrand, n, dt);
        }
        free(integrand);
    }
    free(vavg);
}

/* ------------------------------------------------------------------ aux grid, Cauchy-Green tensor, ridges */

/* flowmap_aux_grid_2D (integration.py:249-464): final positions over the auxiliary stencil
 * (x_i +- h, y_j), (x_i, y_j +- h) [, (x_i, y_j) when eig_main], out[nx, ny, n_aux, 2].
 * Which entries are integrated (everything else stays 0):
 *   eig_main, compute_edge : edge cells (i or j on the border) only the centre point k = 4
 *                            (integration.py:320-343), interior cells all five (345-355)
 *   eig_main, !compute_edge: interior cells only, all five (356-369)
 *   !eig_main              : all four points on [0, nx) x [0, ny) or the interior (425-446) */
int oracle_flowmap_aux_grid_2d(const flow_t *f, double t0, double T, const double *x, int64_t nx,
                               const double *y, int64_t ny, const double *params, double h,
                               int eig_main, int compute_edge, double rtol, double atol,
                               const uint8_t *mask, double *out, int32_t *status, int32_t *steps,
                               int64_t *stats)
{
    const int n_aux = eig_main ? 5 : 4;
    const double aux[5][2] = {{h, 0.0}, {-h, 0.0}, {0.0, h}, {0.0, -h}, {0.0, 0.0}};
    int64_t npts = nx * ny * n_aux;
    double *pts = (double *)malloc(sizeof(double) * 2 * npts);
    uint8_t *m = (uint8_t *)malloc(npts);
    for (int64_t i = 0; i < nx; ++i)
        for (int64_t j = 0; j < ny; ++j) {
            int edge = (i == 0 || i == nx - 1 || j == 0 || j == ny - 1);
            for (int k = 0; k < n_aux; ++k) {
                int64_t q = (i * ny + j) * n_aux + k;
                pts[2 * q] = x[i] + aux[k][0];
                pts[2 * q + 1] = y[j] + aux[k][1];
                int on = !(mask && mask[i * ny + j]);
                if (edge) on = on && compute_edge && (!eig_main || k == 4);
                m[q] = (uint8_t)!on;
            }
        }
    int rc = oracle_flowmap_pts(f, t0, T, pts, npts, params, rtol, atol, m, 2, 1, out, NULL, status,
                                steps, stats);
    free(pts);
    free(m);
    return rc;
}

/* gradF_aux_stencil_2D (utils.py:49-84) */
static inline void grad_aux(const double *fa, int64_t ny, int n_aux, int64_t i, int64_t j, double h,
                            double *g)
{
    const double *c = fa + ((i * ny + j) * n_aux) * 2;
    g[0] = (c[0] - c[2]) / (2 * h);  /* dFxdx */
    g[1] = (c[4] - c[6]) / (2 * h);  /* dFxdy */
    g[2] = (c[1] - c[3]) / (2 * h);  /* dFydx */
    g[3] = (c[5] - c[7]) / (2 * h);  /* dFydy */
}

/* C_tensor_2D (diagnostics.py:68-112): C11, C12, C22 on [2, nx-2) x [2, ny-2) */
void oracle_c_tensor_2d(const double *fm_aux, int64_t nx, int64_t ny, int n_aux, double h,
                        const uint8_t *mask, double *C)
{
    memset(C, 0, sizeof(double) * nx * ny * 3);
#pragma omp parallel for schedule(static)
    for (int64_t i = 2; i < nx - 2; ++i)
        for (int64_t j = 2; j < ny - 2; ++j) {
            if (mask && mask[i * ny + j]) continue;
            double g[4];
            grad_aux(fm_aux, ny, n_aux, i, j, h, g);
            double *c = C + (i * ny + j) * 3;
            c[0] = g[0] * g[0] + g[2] * g[2];
            c[1] = g[0] * g[1] + g[2] * g[3];
            c[2] = g[1] * g[1] + g[3] * g[3];
        }
}

/* np.linalg.eigh / eigvalsh of the symmetric 2x2 matrix [[a, b], [b, c]] as LAPACK computes it
 * (numba and numpy both call ?syevd, uplo 'L'; for n = 2 the tridiagonal reduction is the
 * identity and dsteqr / dsterf either split the matrix when |b| <= sqrt|a| sqrt|c| eps or call
 * dlaev2 / dlae2 once).  dlaev2 is restated from the LAPACK reference documentation; the
 * column / sign conventions were checked against numpy on 20000 random matrices (bit-exact,
 * same signs).  w ascending, v[r][col]: column `col` is the eigenvector of w[col]. */
void oracle_eigh2(double a, double b, double c, double *w, double *v)
{
    const double eps = 1.1102230246251565e-16; /* dlamch('E') */
    if (b == 0.0 || fabs(b) <= (sqrt(fabs(a)) * sqrt(fabs(c))) * eps) {
        if (a <= c) { w[0] = a; w[1] = c; v[0] = 1; v[1] = 0; v[2] = 0; v[3] = 1; }
        else        { w[0] = c; w[1] = a; v[0] = 0; v[1] = 1; v[2] = 1; v[3] = 0; }
        return;
    }
    double sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
    double acmx, acmn, rt, rt1, rt2, cs1, sn1;
    int sgn1, sgn2;
    if (fabs(a) > fabs(c)) { acmx = a; acmn = c; } else { acmx = c; acmn = a; }
    if (adf > ab) { double q = ab / adf; rt = adf * sqrt(1.0 + q * q); }
    else if (adf < ab) { double q = adf / ab; rt = ab * sqrt(1.0 + q * q); }
    else rt = ab * sqrt(2.0);
    if (sm < 0.0) { rt1 = 0.5 * (sm - rt); sgn1 = -1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
    else if (sm > 0.0) { rt1 = 0.5 * (sm + rt); sgn1 = 1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
    else { rt1 = 0.5 * rt; rt2 = -0.5 * rt; sgn1 = 1; }
    double cs;
    if (df >= 0.0) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
    if (fabs(cs) > ab) { double ct = -tb / cs; sn1 = 1.0 / sqrt(1.0 + ct * ct); cs1 = ct * sn1; }
    else if (ab == 0.0) { cs1 = 1.0; sn1 = 0.0; }
    else { double tn = -cs / tb; cs1 = 1.0 / sqrt(1.0 + tn * tn); sn1 = tn * cs1; }
    if (sgn1 == sgn2) { double tn = cs1; cs1 = -sn1; sn1 = tn; }
    /* (cs1, sn1) is the unit eigenvector of rt1, the eigenvalue of larger absolute value */
    if (rt1 >= rt2) { w[0] = rt2; w[1] = rt1; v[0] = -sn1; v[1] = cs1; v[2] = cs1; v[3] = sn1; }
    else            { w[0] = rt1; w[1] = rt2; v[0] = cs1; v[1] = -sn1; v[2] = sn1; v[3] = cs1; }
}

/* C_eig_2D (diagnostics.py:200-244): eigvals[nx,ny,2] ascending, eigvecs[nx,ny,2,2] */
void oracle_c_eig_2d(const double *fm, int64_t nx, int64_t ny, double dx, double dy,
                     const uint8_t *mask, double *eigvals, double *eigvecs)
{
    memset(eigvals, 0, sizeof(double) * nx * ny * 2);
    memset(eigvecs, 0, sizeof(double) * nx * ny * 4);
#pragma omp parallel for schedule(static)
    for (int64_t i = 1; i < nx - 1; ++i)
        for (int64_t j = 1; j < ny - 1; ++j) {
            if (mask && mask[i * ny + j]) continue;
#define F(I, J, C) fm[(((I)*ny) + (J)) * 2 + (C)]
            double dxdx = (F(i + 1, j, 0) - F(i - 1, j, 0)) / (2 * dx);
            double dxdy = (F(i, j + 1, 0) - F(i, j - 1, 0)) / (2 * dy);
            double dydx = (F(i + 1, j, 1) - F(i - 1, j, 1)) / (2 * dx);
            double dydy = (F(i, j + 1, 1) - F(i, j - 1, 1)) / (2 * dy);
#undef F
            double off = dxdx * dxdy + dydx * dydy;
            oracle_eigh2(dxdx * dxdx + dydx * dydx, off, dxdy * dxdy + dydy * dydy,
                         eigvals + (i * ny + j) * 2, eigvecs + (i * ny + j) * 4);
        }
}

/* C_eig_aux_2D (diagnostics.py:115-197): eig_main -> eigenvalues of the main-grid tensor
 * (gradF_main_stencil_2D on the centre points, utils.py:87-124), eigenvectors of the aux-grid
 * tensor, on [2, nx-2) x [2, ny-2); otherwise both from the aux grid on [1, nx-1) x [1, ny-1) */
void oracle_c_eig_aux_2d(const double *fm_aux, int64_t nx, int64_t ny, int n_aux, double dx,
                         double dy, double h, int eig_main, const uint8_t *mask, double *eigvals,
                         double *eigvecs)
{
    memset(eigvals, 0, sizeof(double) * nx * ny * 2);
    memset(eigvecs, 0, sizeof(double) * nx * ny * 4);
    int64_t lo = eig_main ? 2 : 1;
#pragma omp parallel for schedule(static)
    for (int64_t i = lo; i < nx - lo; ++i)
        for (int64_t j = lo; j < ny - lo; ++j) {
            if (mask && mask[i * ny + j]) continue;
            double g[4], w[2], v[4];
            grad_aux(fm_aux, ny, n_aux, i, j, h, g);
            oracle_eigh2(g[0] * g[0] + g[2] * g[2], g[0] * g[1] + g[2] * g[3],
                         g[1] * g[1] + g[3] * g[3], w, v);
            if (eig_main) {
#define FA(I, J, C) fm_aux[((((I)*ny) + (J)) * n_aux + (n_aux - 1)) * 2 + (C)]
                double dxdx = (FA(i + 1, j, 0) - FA(i - 1, j, 0)) / (2 * dx);
                double dxdy = (FA(i, j + 1, 0) - FA(i, j - 1, 0)) / (2 * dy);
                double dydx = (FA(i + 1, j, 1) - FA(i - 1, j, 1)) / (2 * dx);
                double dydy = (FA(i, j + 1, 1) - FA(i, j - 1, 1)) / (2 * dy);
#undef FA
                double vm[4];
                oracle_eigh2(dxdx * dxdx + dydx * dydx, dxdx * dxdy + dydx * dydy,
                             dxdy * dxdy + dydy * dydy, w, vm);
            }
            for (int k = 0; k < 2; ++k) eigvals[(i * ny + j) * 2 + k] = w[k];
            for (int k = 0; k < 4; ++k) eigvecs[(i * ny + j) * 4 + k] = v[k];
        }
}

/* ftle_from_eig (diagnostics.py:247-269) */
void oracle_ftle_from_eig(const double *eigval_max, int64_t n, double T, double *ftle)
{
    for (int64_t q = 0; q < n; ++q)
        ftle[q] = (eigval_max[q] > 1) ? log(eigval_max[q]) / (2 * fabs(T)) : 0.0;
}

/* _ftle_ridge_pts_connect (extraction/ridges.py:232-318), which is ftle_ridge_pts (9-76) plus the
 * per-pixel eigenvector and second directional derivative: r_pts[nx*ny, 3] (-1 where no ridge
 * point), r_vec[nx*ny, 2], sdd[nx*ny].  f_min is 0 or np.percentile(f, percentile), computed by
 * the caller.  Returns the number of ridge points. */
int64_t oracle_ftle_ridge_pts(const double *f, const double *evec, int64_t nx, int64_t ny,
                              const double *x, const double *y, double dx, double dy,
                              double sdd_thresh, double f_min, double *r_pts, double *r_vec,
                              double *sdd)
{
    /* dx = x[1] - x[0], dy = y[1] - y[0] (ridges.py:39-40), passed by the caller so that a row
     * slab of a larger grid can use the spacing of the full grid */
    for (int64_t q = 0; q < nx * ny; ++q) {
        r_pts[3 * q] = r_pts[3 * q + 1] = r_pts[3 * q + 2] = -1.0;
        r_vec[2 * q] = r_vec[2 * q + 1] = 0.0;
        sdd[q] = 0.0;
    }
    int64_t count = 0;
#pragma omp parallel for schedule(static) reduction(+ : count)
    for (int64_t i = 2; i < nx - 2; ++i)
        for (int64_t j = 2; j < ny - 2; ++j) {
#define F(I, J) f[(I)*ny + (J)]
            double f0 = F(i, j);
            if (!(f0 > f_min)) continue;
            double fx = (F(i + 1, j) - F(i - 1, j)) / (2 * dx);
            double fy = (F(i, j + 1) - F(i, j - 1)) / (2 * dy);
            double fxx = (F(i + 1, j) - 2 * F(i, j) + F(i - 1, j)) / (dx * dx);
            double fyy = (F(i, j + 1) - 2 * F(i, j) + F(i, j - 1)) / (dy * dy);
            double fxy = (F(i + 1, j + 1) - F(i + 1, j - 1) - F(i - 1, j + 1) + F(i - 1, j - 1)) /
                         (4 * dx * dy);
#undef F
            double ex = evec[(i * ny + j) * 2], ey = evec[(i * ny + j) * 2 + 1];
            double c2 = ex * (fxx * ex + fxy * ey) + ey * (fxy * ex + fyy * ey);
            if (c2 < -sdd_thresh) {
                double t = -(fx * ex + fy * ey) / c2;
                if (fabs(t * ex) <= dx / 2 && fabs(t * ey) <= dy / 2) {
                    int64_t k = i * ny + j;
                    r_pts[3 * k] = x[i] + t * ex;
                    r_pts[3 * k + 1] = y[j] + t * ey;
                    r_vec[2 * k] = ex;
                    r_vec[2 * k + 1] = ey;
                    sdd[k] = c2;
                    ++count;
                }
            }
        }
    return count;
}

/* ------------------------------------------------------------------ flow-map composition */

/* eval_linear(grid, F, pts, xto.CONSTANT) on a 2-D grid ((a0,b0,n0),(a1,b1,n1)); F has element
 * stride `st` (flowmaps[k, :, :, c] is a strided view).  Outside the grid -> 0. */
static double eval_linear2(const double *g6, const double *F, int64_t st, double p0, double p1)
{
    double a0 = g6[0], b0 = g6[1], a1 = g6[3], b1 = g6[4];
    int64_t n0 = (int64_t)g6[2], n1 = (int64_t)g6[5];
    if (p0 < a0 || p0 > b0 || p1 < a1 || p1 > b1) return 0.0;
    int64_t i0, i1;
    double l0, l1;
    axis_locate(a0, b0, n0, p0, &i0, &l0);
    axis_locate(a1, b1, n1, p1, &i1, &l1);
    const double *c = F + (i0 * n1 + i1) * st;
    double v = 0.0;
    for (int a = 0; a < 2; ++a) {
        double wa = a ? l0 : 1.0 - l0;
        const double *cc = c + a * n1 * st;
        v += wa * ((1.0 - l1) * cc[0] + l1 * cc[st]);
    }
    return v;
}

/* flowmap_composition (integration.py:609-644): flowmaps[nT, nx, ny, 2] -> composed[nx, ny, 2].
 * pts = flowmaps[0]; for k in 1..nT-2: pts = flowmaps[k](pts); composed = flowmaps[nT-1](pts)
 * (so nT == 1 applies flowmaps[0] to itself, as the reference does). */
void oracle_flowmap_composition(const double *flowmaps, const double *grid6, int64_t nT,
                                double *composed)
{
    int64_t nx = (int64_t)grid6[2], ny = (int64_t)grid6[5], np = nx * ny;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < np; ++q) {
        double px = flowmaps[2 * q], py = flowmaps[2 * q + 1];
        for (int64_t k = 1; k < nT - 1; ++k) {
            const double *F = flowmaps + k * np * 2;
            double fx = eval_linear2(grid6, F, 2, px, py);
            double fy = eval_linear2(grid6, F + 1, 2, px, py);
            px = fx;
            py = fy;
        }
        const double *F = flowmaps + (nT - 1) * np * 2;
        composed[2 * q] = eval_linear2(grid6, F, 2, px, py);
        composed[2 * q + 1] = eval_linear2(grid6, F + 1, 2, px, py);
    }
}

/* ------------------------------------------------------------------ helpers ------------ */

flow_t *oracle_flow_new(int kind)
{
    flow_t *f = (flow_t *)calloc(1, sizeof(flow_t));
    f->kind = kind;
    f->ndim = (kind == FLOW_ABC) ? 3 : 2;
    f->r = 6371.0;
    return f;
}

static void fill_spline(spline3_t *s, const double *grid9, const double *C, int extrap)
{
    for (int d = 0; d < 3; ++d) {
        s->a[d] = grid9[3 * d];
        s->b[d] = grid9[3 * d + 1];
        s->n[d] = (int64_t)grid9[3 * d + 2];
    }
    s->C = C;
    s->extrap = extrap;
}

/* get_flow_2D (flows.py:121-258); coefficient arrays are borrowed, caller keeps them alive */
flow_t *oracle_flow_new_spline(const double *grid9, const double *Cu, const double *Cv,
                               int spherical, int extrap, double r)
{
    flow_t *f = oracle_flow_new(FLOW_SPLINE2D);
    f->spherical = spherical;
    f->r = r;
    fill_spline(&f->u, grid9, Cu, extrap);
    fill_spline(&f->v, grid9, Cv, extrap);
    return f;
}

/* get_flow_linear_2D (flows.py:418-506): raw (nt, nx, ny) arrays, trilinear */
flow_t *oracle_flow_new_linear(const double *grid9, const double *U, const double *V, int spherical,
                               int extrap, double r)
{
    flow_t *f = oracle_flow_new_spline(grid9, U, V, spherical, extrap, r);
    f->kind = FLOW_LINEAR2D;
    return f;
}

spline3_t *oracle_scalar_new(const double *grid9, const double *C, int extrap)
{
    spline3_t *s = (spline3_t *)calloc(1, sizeof(spline3_t));
    fill_spline(s, grid9, C, extrap);
    return s;
}

void oracle_free(void *p) { free(p); }

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
