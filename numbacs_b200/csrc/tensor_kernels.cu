// tensor_kernels.cu -- Cauchy-Green tensor, its eigen-pairs, FTLE-from-eigenvalue and FTLE ridge
// points: the consumers of the flow map next to ftle_grid_2D.
//
// Replaces (paths relative to /root/reference/src/numbacs):
//   C_tensor_2D            diagnostics.py:68-112   (gradF_aux_stencil_2D, utils.py:49-84)
//   C_eig_aux_2D           diagnostics.py:115-197  (+ gradF_main_stencil_2D, utils.py:87-124)
//   C_eig_2D               diagnostics.py:200-244  (gradF_stencil_2D, utils.py:9-46)
//   ftle_from_eig          diagnostics.py:247-269
//   ftle_ridge_pts         extraction/ridges.py:9-76
//   _ftle_ridge_pts_connect extraction/ridges.py:232-318
//   np.percentile(f, p)    (ridges.py:45, 279) -> two order statistics by radix select
//
// All of these are HBM-bound streaming passes (16-80 B read, 8-48 B written per pixel), far below
// the FP64 budget, so the arithmetic is written with explicitly rounded operations
// (__dadd_rn / __dmul_rn / __ddiv_rn / __dsqrt_rn: no FMA contraction) in the reference's
// operation order: eigenvalues, eigenvector components AND signs, and ridge points come out
// bit-identical to numpy/numba + LAPACK, which is what makes `eigvecs[..., 1]` safe to feed to
// code that was developed against the reference.
//
// np.linalg.eigh on a symmetric 2x2 (numba and numpy: LAPACK ?syevd, uplo 'L'): the tridiagonal
// reduction is the identity for n = 2; dsteqr / dsterf then either split the matrix when the
// off-diagonal is negligible (|b| <= sqrt|a| sqrt|c| eps, eigenvectors = unit vectors, sorted) or
// call dlaev2 / dlae2 once.  eigh2() below restates that; the oracle holds the same restatement
// and is pinned against the live reference (tests/test_oracle_tensor_golden.py).
#include <cfloat>

#include "common.cuh"
#include "launch.cuh"

namespace b200cs {

namespace {

__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

// a / c for a divisor known on the host, with rc = RN(1/c) precomputed there: q = RN(a rc), the
// exact residual r = a - q c (one FMA), q' = RN(q + r rc).  The result is the correctly rounded
// quotient (Markstein; the residual correction leaves an error of ~2^-52 ulp before the final
// rounding, so only quotients within that distance of a rounding midpoint could differ -- none in
// 1e8 random trials against a true division) in three FP64 instructions instead of the ~20 of the
// generic IEEE division sequence.  Operands here are O(1) physical quantities: no under/overflow.
struct Divisor {
    double c, rc;
};
__device__ __forceinline__ double dvd(double a, const Divisor &d) {
    const double q = mul(a, d.rc);
    return fma(fma(-q, d.c, a), d.rc, q);
}
inline Divisor make_divisor(double c) { return Divisor{c, 1.0 / c}; }

// w ascending; v = [v00, v01, v10, v11], column c is the eigenvector of w[c]
__device__ __forceinline__ void eigh2(double a, double b, double c, double (&w)[2], double (&v)[4]) {
    const double eps = 1.1102230246251565e-16;  // dlamch('E')
    // the split test |b| <= sqrt|a| sqrt|c| eps costs two square roots; it can only hold when
    // b^2 <= |a c| eps^2 (1 + a few ulps), so a product test with a safety margin screens it out
    // for practically every pixel (a product that underflows to 0 falls through to the exact test
    // or, with b^2 > 0, is correctly classified: |b| > 1e-162 > sqrt(1e-308) eps)
    const bool maybe_split = !(mul(b, b) > mul(fabs(mul(a, c)), 1.2325951644078310e-32 * 1.000001));
    if (b == 0.0 ||
        (maybe_split && fabs(b) <= mul(mul(__dsqrt_rn(fabs(a)), __dsqrt_rn(fabs(c))), eps))) {
        const bool keep = a <= c;
        w[0] = keep ? a : c;
        w[1] = keep ? c : a;
        v[0] = keep ? 1.0 : 0.0;
        v[1] = keep ? 0.0 : 1.0;
        v[2] = v[1];
        v[3] = v[0];
        return;
    }
    // dlaev2(a, b, c)
    const double sm = add(a, c), df = sub(a, c), adf = fabs(df), tb = add(b, b), ab = fabs(tb);
    const bool a_big = fabs(a) > fabs(c);
    const double acmx = a_big ? a : c, acmn = a_big ? c : a;
    double rt;
    if (adf > ab) {
        const double q = dvd(ab, adf);
        rt = mul(adf, __dsqrt_rn(add(1.0, mul(q, q))));
    } else if (adf < ab) {
        const double q = dvd(adf, ab);
        rt = mul(ab, __dsqrt_rn(add(1.0, mul(q, q))));
    } else {
        rt = mul(ab, __dsqrt_rn(2.0));
    }
    double rt1, rt2;
    int sgn1;
    if (sm < 0.0) {
        rt1 = mul(0.5, sub(sm, rt));
        sgn1 = -1;
        rt2 = sub(mul(dvd(acmx, rt1), acmn), mul(dvd(b, rt1), b));
    } else if (sm > 0.0) {
        rt1 = mul(0.5, add(sm, rt));
        sgn1 = 1;
        rt2 = sub(mul(dvd(acmx, rt1), acmn), mul(dvd(b, rt1), b));
    } else {
        rt1 = mul(0.5, rt);
        rt2 = mul(-0.5, rt);
        sgn1 = 1;
    }
    const int sgn2 = (df >= 0.0) ? 1 : -1;
    const double cs = (df >= 0.0) ? add(df, rt) : sub(df, rt);
    double cs1, sn1;
    if (fabs(cs) > ab) {
        const double ct = dvd(-tb, cs);
        sn1 = dvd(1.0, __dsqrt_rn(add(1.0, mul(ct, ct))));
        cs1 = mul(ct, sn1);
    } else if (ab == 0.0) {
        cs1 = 1.0;
        sn1 = 0.0;
    } else {
        const double tn = dvd(-cs, tb);
        cs1 = dvd(1.0, __dsqrt_rn(add(1.0, mul(tn, tn))));
        sn1 = mul(tn, cs1);
    }
    if (sgn1 == sgn2) {
        const double tn = cs1;
        cs1 = -sn1;
        sn1 = tn;
    }
    // (cs1, sn1): unit eigenvector of rt1, the eigenvalue of larger absolute value
    if (rt1 >= rt2) {
        w[0] = rt2; w[1] = rt1;
        v[0] = -sn1; v[1] = cs1; v[2] = cs1; v[3] = sn1;
    } else {
        w[0] = rt1; w[1] = rt2;
        v[0] = cs1; v[1] = -sn1; v[2] = sn1; v[3] = cs1;
    }
}

// C11, C12, C22 from the four gradient entries (diagnostics.py:101-107, 152-167, 236-238)
__device__ __forceinline__ void cg_tensor(double dxdx, double dxdy, double dydx, double dydy, double &c11,
                                          double &c12, double &c22) {
    c11 = add(mul(dxdx, dxdx), mul(dydx, dydx));
    c12 = add(mul(dxdx, dxdy), mul(dydx, dydy));
    c22 = add(mul(dxdy, dxdy), mul(dydy, dydy));
}

// gradF_aux_stencil_2D (utils.py:80-84): cell = fm_aux + ((i*ny + j)*n_aux)*2
__device__ __forceinline__ void grad_aux(const double *__restrict__ cell, const Divisor &two_h, double &dxdx,
                                         double &dxdy, double &dydx, double &dydy) {
    const double2 p0 = __ldg(reinterpret_cast<const double2 *>(cell));
    const double2 p1 = __ldg(reinterpret_cast<const double2 *>(cell) + 1);
    const double2 p2 = __ldg(reinterpret_cast<const double2 *>(cell) + 2);
    const double2 p3 = __ldg(reinterpret_cast<const double2 *>(cell) + 3);
    dxdx = dvd(sub(p0.x, p1.x), two_h);
    dxdy = dvd(sub(p2.x, p3.x), two_h);
    dydx = dvd(sub(p0.y, p1.y), two_h);
    dydy = dvd(sub(p2.y, p3.y), two_h);
}

constexpr int kTB = 256;

// ---- C_tensor_2D -------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTB)
c_tensor_kernel(const double *__restrict__ fa, long long nx, long long ny, int n_aux, Divisor two_h,
                const uint8_t *__restrict__ mask, double *__restrict__ C) {
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    if (q >= nx * ny) return;
    const long long i = q / ny, j = q - i * ny;
    double c11 = 0.0, c12 = 0.0, c22 = 0.0;
    if (i >= 2 && i < nx - 2 && j >= 2 && j < ny - 2 && !(mask && mask[q])) {
        double dxdx, dxdy, dydx, dydy;
        grad_aux(fa + q * n_aux * 2, two_h, dxdx, dxdy, dydx, dydy);
        cg_tensor(dxdx, dxdy, dydx, dydy, c11, c12, c22);
    }
    C[3 * q] = c11;
    C[3 * q + 1] = c12;
    C[3 * q + 2] = c22;
}

// ---- C_eig_2D / C_eig_aux_2D --------------------------------------------------------------------
// fm is [nx, ny, n_aux, 2]; the main-grid stencil (gradF_stencil_2D / gradF_main_stencil_2D) reads
// the LAST of the n_aux points of the four neighbours, the aux stencil the first four points of
// the cell itself.  `lo`: first interior index (2 when both stencils are used, else 1).
__global__ void __launch_bounds__(kTB)
c_eig_kernel(const double *__restrict__ fm, long long nx, long long ny, int n_aux, Divisor two_h, Divisor two_dx,
             Divisor two_dy, int aux_vecs, int main_vals, int lo, const uint8_t *__restrict__ mask,
             double *__restrict__ eigvals, double *__restrict__ eigvecs, double *__restrict__ ftle,
             double two_absT) {
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    if (q >= nx * ny) return;
    const long long i = q / ny, j = q - i * ny;
    double w[2] = {0.0, 0.0}, v[4] = {0.0, 0.0, 0.0, 0.0};
    if (i >= lo && i < nx - lo && j >= lo && j < ny - lo && !(mask && mask[q])) {
        double c11, c12, c22;
        if (aux_vecs) {
            double dxdx, dxdy, dydx, dydy;
            grad_aux(fm + q * n_aux * 2, two_h, dxdx, dxdy, dydx, dydy);
            cg_tensor(dxdx, dxdy, dydx, dydy, c11, c12, c22);
            eigh2(c11, c12, c22, w, v);
        }
        if (main_vals) {
            const long long rs = ny * n_aux;  // double2 elements per grid row
            const double2 *c = reinterpret_cast<const double2 *>(fm) + q * n_aux + (n_aux - 1);
            const double2 up = __ldg(c + rs), dn = __ldg(c - rs), rt = __ldg(c + n_aux), lf = __ldg(c - n_aux);
            const double dxdx = dvd(sub(up.x, dn.x), two_dx);
            const double dxdy = dvd(sub(rt.x, lf.x), two_dy);
            const double dydx = dvd(sub(up.y, dn.y), two_dx);
            const double dydy = dvd(sub(rt.y, lf.y), two_dy);
            cg_tensor(dxdx, dxdy, dydx, dydy, c11, c12, c22);
            if (aux_vecs) {
                double vm[4];
                eigh2(c11, c12, c22, w, vm);  // eigenvalues from the main grid, vectors from the aux grid
            } else {
                eigh2(c11, c12, c22, w, v);
            }
        }
    }
    reinterpret_cast<double2 *>(eigvals)[q] = make_double2(w[0], w[1]);
    reinterpret_cast<double2 *>(eigvecs)[2 * q] = make_double2(v[0], v[1]);
    reinterpret_cast<double2 *>(eigvecs)[2 * q + 1] = make_double2(v[2], v[3]);
    // ftle_from_eig fused (the same operations as ftle_from_eig_kernel on eigvals[..., 1]): saves the
    // separate pass that re-reads what was just written
    if (ftle) ftle[q] = (w[1] > 1.0) ? dvd(log(w[1]), two_absT) : 0.0;
}

// ---- ftle_from_eig ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTB)
ftle_from_eig_kernel(const double *__restrict__ e, long long n, long long stride, double two_absT,
                     double *__restrict__ out) {
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    if (q >= n) return;
    const double lam = __ldg(e + q * stride);
    out[q] = (lam > 1.0) ? dvd(log(lam), two_absT) : 0.0;
}

// ---- ridge points ---------------------------------------------------------------------------------
struct RidgeArgs {
    const double *f, *ev, *x, *y;
    long long ev_ps, ev_cs, nx, ny;
    Divisor two_dx, two_dy, dx2, dy2, four_dxdy;
    double half_dx, half_dy, sdd_thresh, f_min;
};

// the per-pixel test of ridges.py:49-76 / 287-316; true when (i, j) carries a ridge point
__device__ __forceinline__ bool ridge_at(const RidgeArgs &R, long long q, double &px, double &py, double &ex,
                                         double &ey, double &c2) {
    const long long i = q / R.ny, j = q - i * R.ny;
    if (i < 2 || i >= R.nx - 2 || j < 2 || j >= R.ny - 2) return false;
    const double *c = R.f + q;
    const double f0 = __ldg(c);
    if (!(f0 > R.f_min)) return false;
    const long long ny = R.ny;
    const double fu = __ldg(c + ny), fd = __ldg(c - ny), fr = __ldg(c + 1), fl = __ldg(c - 1);
    const double fx = dvd(sub(fu, fd), R.two_dx);
    const double fy = dvd(sub(fr, fl), R.two_dy);
    const double two_f0 = mul(2.0, f0);
    const double fxx = dvd(add(sub(fu, two_f0), fd), R.dx2);
    const double fyy = dvd(add(sub(fr, two_f0), fl), R.dy2);
    const double fxy = dvd(add(sub(sub(__ldg(c + ny + 1), __ldg(c + ny - 1)), __ldg(c - ny + 1)), __ldg(c - ny - 1)),
                           R.four_dxdy);
    ex = __ldg(R.ev + q * R.ev_ps);
    ey = __ldg(R.ev + q * R.ev_ps + R.ev_cs);
    c2 = add(mul(ex, add(mul(fxx, ex), mul(fxy, ey))), mul(ey, add(mul(fxy, ex), mul(fyy, ey))));
    if (!(c2 < -R.sdd_thresh)) return false;
    const double t = dvd(-add(mul(fx, ex), mul(fy, ey)), c2);
    const double tx = mul(t, ex), ty = mul(t, ey);
    if (!(fabs(tx) <= R.half_dx && fabs(ty) <= R.half_dy)) return false;
    px = add(__ldg(R.x + i), tx);
    py = add(__ldg(R.y + j), ty);
    return true;
}

// pass 1: full per-pixel arrays (optional) + ridge points per block
__global__ void __launch_bounds__(kTB)
ridge_detect_kernel(const __grid_constant__ RidgeArgs R, double *__restrict__ r_pts, double *__restrict__ r_vec,
                    double *__restrict__ sdd, int *__restrict__ block_counts) {
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    bool hit = false;
    double px = -1.0, py = -1.0, ex = 0.0, ey = 0.0, c2 = 0.0;
    if (q < R.nx * R.ny) {
        hit = ridge_at(R, q, px, py, ex, ey, c2);
        if (r_pts) {
            r_pts[3 * q] = hit ? px : -1.0;
            r_pts[3 * q + 1] = hit ? py : -1.0;
            r_pts[3 * q + 2] = -1.0;  // ridge-number placeholder (ridges.py:271)
        }
        if (r_vec) {
            r_vec[2 * q] = hit ? ex : 0.0;
            r_vec[2 * q + 1] = hit ? ey : 0.0;
        }
        if (sdd) sdd[q] = hit ? c2 : 0.0;
    }
    const int cnt = __syncthreads_count(hit ? 1 : 0);
    if (block_counts && threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

// pass 2: exclusive scan of the block counts (one block; nblocks is at most a few million, every
// thread takes four consecutive counts per round so a million counts are 256 rounds)
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int *__restrict__ counts, long long nblocks, long long *__restrict__ offsets,
                   long long *__restrict__ total) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long base = 0; base < nblocks; base += 4096) {
        const long long idx = base + 4LL * threadIdx.x;
        int c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = (idx + k < nblocks) ? counts[idx + k] : 0;
        const long long v = (long long)c[0] + c[1] + c[2] + c[3];
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            long long ws = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sums[lane] = ws;  // inclusive over warps
        }
        __syncthreads();
        const long long carry = carry_s;
        long long run = carry + ((wid > 0) ? warp_sums[wid - 1] : 0) + inc - v;  // exclusive prefix of c[0]
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (idx + k < nblocks) offsets[idx + k] = run;
            run += c[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

// pass 3: ridge points in raveled-index order (r_pts[ridge_bool, :], ridges.py:76)
__global__ void __launch_bounds__(kTB)
ridge_compact_kernel(const __grid_constant__ RidgeArgs R, const long long *__restrict__ offsets,
                     double *__restrict__ pts, long long capacity) {
    __shared__ int warp_cnt[kTB / 32];
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    bool hit = false;
    double px = 0.0, py = 0.0, ex, ey, c2;
    if (q < R.nx * R.ny) hit = ridge_at(R, q, px, py, ex, ey, c2);
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    if (!hit) return;
    long long pos = offsets[blockIdx.x] + __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < wid; ++w) pos += warp_cnt[w];
    if (pos < capacity) {
        pts[2 * pos] = px;
        pts[2 * pos + 1] = py;
    }
}

// ---- ridge points, compact output only: one evaluation per pixel, hits kept as a bit mask -------
// (round 2) The three-pass form above evaluates the ridge test twice for every pixel (detect,
// compact) and scans one count per 128 pixels in a single block.  Here a block owns kRB = 2048
// consecutive pixels: pass 1 evaluates the test once and leaves one ballot word per warp and
// iteration (1 bit per pixel) plus the block's count; the scan runs over 16x fewer counts; pass 3
// reads the 64 words of its block, turns them into positions with popc prefix sums and re-evaluates
// the test only for the set bits (0.6 % of the pixels of the 16384^2 double-gyre field).
constexpr int kRB = 2048;                 // pixels per block
constexpr int kRIter = kRB / kTB;         // iterations per thread

__global__ void __launch_bounds__(kTB)
ridge_detect_bits_kernel(const __grid_constant__ RidgeArgs R, unsigned *__restrict__ bits, int *__restrict__ block_counts) {
    const long long base = (long long)blockIdx.x * kRB;
    const long long np = R.nx * R.ny;
    int mine = 0;
#pragma unroll 4
    for (int k = 0; k < kRIter; ++k) {
        const long long q = base + (long long)k * kTB + threadIdx.x;
        bool hit = false;
        double px, py, ex, ey, c2;
        if (q < np) hit = ridge_at(R, q, px, py, ex, ey, c2);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if ((threadIdx.x & 31) == 0) {
            bits[(base + (long long)k * kTB + threadIdx.x) >> 5] = bal;
            mine += __popc(bal);
        }
    }
    __shared__ int wsum[kTB / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < kTB / 32; ++w) tot += wsum[w];
        block_counts[blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(kTB)
ridge_compact_bits_kernel(const __grid_constant__ RidgeArgs R, const unsigned *__restrict__ bits,
                          const long long *__restrict__ offsets, double *__restrict__ pts, long long capacity) {
    constexpr int kWords = kRB / 32;                       // 64 ballot words per block, in raveled order
    __shared__ unsigned words[kWords];
    __shared__ int wpre[kWords];
    const long long base = (long long)blockIdx.x * kRB;
    if (threadIdx.x < kWords) words[threadIdx.x] = bits[(base >> 5) + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < kWords; ++w) {
            wpre[w] = run;
            run += __popc(words[w]);
        }
    }
    __syncthreads();
    const long long off = offsets[blockIdx.x];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < kRIter; ++k) {
        const int w = k * (kTB / 32) + wid;                // word of pixels base + k*kTB + wid*32 ..
        const unsigned word = words[w];
        if (!((word >> lane) & 1u)) continue;
        const long long q = base + (long long)k * kTB + threadIdx.x;
        double px = 0.0, py = 0.0, ex, ey, c2;
        ridge_at(R, q, px, py, ex, ey, c2);
        const long long pos = off + wpre[w] + __popc(word & ((1u << lane) - 1u));
        if (pos < capacity) {
            pts[2 * pos] = px;
            pts[2 * pos + 1] = py;
        }
    }
}

// ---- connected ridges (ftle_ridges, ridges.py:175-229) ------------------------------------------
// The reference labels the 8-connected components of ridge_bool with scipy.ndimage.label and then
// gathers each component's points with one full-grid comparison per label.  Here: union-find over
// the ridge pixels (every pixel links to its W / NW / N / NE ridge neighbours, roots are the
// smallest raveled index of their component, so sorting roots reproduces scipy's raster-order
// label numbering), then the same ordered compaction as above carrying each point's root.
__global__ void __launch_bounds__(kTB)
ridge_flag_kernel(const __grid_constant__ RidgeArgs R, int *__restrict__ parent, int *__restrict__ block_counts) {
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    bool hit = false;
    if (q < R.nx * R.ny) {
        double px, py, ex, ey, c2;
        hit = ridge_at(R, q, px, py, ex, ey, c2);
        parent[q] = hit ? (int)q : -1;
    }
    const int cnt = __syncthreads_count(hit ? 1 : 0);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

// parent pointers only ever decrease and always stay inside the component, so a stale value is
// still a valid ancestor; the loads go to L2 (where the atomics land) to keep the walks short
__device__ __forceinline__ int uf_find(const int *parent, int x) {
    int p = __ldcg(parent + x);
    while (p != x) {
        x = p;
        p = __ldcg(parent + x);
    }
    return x;
}

__device__ __forceinline__ void uf_union(int *parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&parent[b], a);  // link the larger root under the smaller
        if (old == b) return;
        b = old;  // somebody re-rooted b meanwhile: merge what it points to now
    }
}

__global__ void __launch_bounds__(kTB)
ridge_merge_kernel(int *parent, long long nx, long long ny) {
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    if (q >= nx * ny || parent[q] < 0) return;
    const long long i = q / ny, j = q - i * ny;
    if (j > 0 && parent[q - 1] >= 0) uf_union(parent, (int)q, (int)(q - 1));
    if (i > 0) {
        const long long u = q - ny;
        if (parent[u] >= 0) uf_union(parent, (int)q, (int)u);
        if (j > 0 && parent[u - 1] >= 0) uf_union(parent, (int)q, (int)(u - 1));
        if (j < ny - 1 && parent[u + 1] >= 0) uf_union(parent, (int)q, (int)(u + 1));
    }
}

__global__ void __launch_bounds__(kTB)
ridge_compact_roots_kernel(const __grid_constant__ RidgeArgs R, const int *__restrict__ parent,
                           const long long *__restrict__ offsets, double *__restrict__ pts,
                           long long *__restrict__ roots, long long capacity) {
    __shared__ int warp_cnt[kTB / 32];
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    const bool hit = (q < R.nx * R.ny) && parent[q] >= 0;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    if (!hit) return;
    long long pos = offsets[blockIdx.x] + __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < wid; ++w) pos += warp_cnt[w];
    if (pos < capacity) {
        double px, py, ex, ey, c2;
        ridge_at(R, q, px, py, ex, ey, c2);
        pts[2 * pos] = px;
        pts[2 * pos + 1] = py;
        roots[pos] = uf_find(parent, (int)q);
    }
}

// ---- flow-map composition (flowmap_composition, integration.py:609-644) --------------------------
// composed = F_{nT-1}( ... F_2(F_1(F_0)) ... ): every grid point walks through the chain of
// intermediate flow maps by bilinear interpolation (interpolation.splines.eval_linear on the 2-D
// grid, CONSTANT extrapolation: a component evaluated outside the grid is 0).  The reference makes
// 2 (nT-1) full-grid eval_linear passes through a points array; here ONE kernel keeps the running
// position in registers and gathers the four (x, y) taps of each map as double2 -- 16 B read +
// 16 B written per point of HBM traffic, the (nT-1) x 64 B of gathers are L2 hits (a flow map is
// smooth: neighbouring threads hit neighbouring cells).
struct Grid2 {
    double a[2], b[2], delta[2], inv_delta[2];
    int n[2];
};

__device__ __forceinline__ void locate2(const Grid2 &g, int d, double x, int &i, double &lam) {
    const double dd = x - g.a[d];
    double fi = floor(dd * g.inv_delta[d]);
    fi = fmin(fmax(fi, 0.0), (double)(g.n[d] - 2));
    i = (int)fi;
    lam = __dsub_rn(dd, __dmul_rn(fi, g.delta[d])) * g.inv_delta[d];
}

__device__ __forceinline__ double2 bilinear2(const Grid2 &g, const double2 *__restrict__ F, double px, double py) {
    if (px < g.a[0] || px > g.b[0] || py < g.a[1] || py > g.b[1]) return make_double2(0.0, 0.0);
    int i0, i1;
    double l0, l1;
    locate2(g, 0, px, i0, l0);
    locate2(g, 1, py, i1, l1);
    const double2 *c = F + (long long)i0 * g.n[1] + i1;
    const double2 c00 = __ldg(c), c01 = __ldg(c + 1), c10 = __ldg(c + g.n[1]), c11 = __ldg(c + g.n[1] + 1);
    const double m1 = 1.0 - l1, m0 = 1.0 - l0;
    double2 v;
    v.x = fma(l0, fma(l1, c11.x, m1 * c10.x), m0 * fma(l1, c01.x, m1 * c00.x));
    v.y = fma(l0, fma(l1, c11.y, m1 * c10.y), m0 * fma(l1, c01.y, m1 * c00.y));
    return v;
}

__global__ void __launch_bounds__(kTB)
composition_kernel(const double2 *__restrict__ fms, const __grid_constant__ Grid2 g, long long nT,
                   double2 *__restrict__ out) {
    const long long np = (long long)g.n[0] * g.n[1];
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    if (q >= np) return;
    // blockIdx.y = frame of a sliding window: frame f composes maps f .. f + nT - 1
    fms += (long long)blockIdx.y * np;
    out += (long long)blockIdx.y * np;
    double2 p = __ldg(fms + q);
    for (long long k = 1; k < nT - 1; ++k) p = bilinear2(g, fms + k * np, p.x, p.y);
    out[q] = bilinear2(g, fms + (nT - 1) * np, p.x, p.y);
}

// ---- binary_mask_dilation (utils.py:1923-1985) ----------------------------------------------------
// one byte per pixel: a pixel is set when it or one of its 4 (corners: 8) neighbours is set
__global__ void __launch_bounds__(kTB)
mask_dilation_kernel(const uint8_t *__restrict__ mask, long long nx, long long ny, int corners,
                     uint8_t *__restrict__ out) {
    const long long q = (long long)blockIdx.x * kTB + threadIdx.x;
    if (q >= nx * ny) return;
    const long long i = q / ny, j = q - i * ny;
    const bool up = i > 0, dn = i < nx - 1, lf = j > 0, rt = j < ny - 1;
    bool v = mask[q] != 0;
    v = v || (up && mask[q - ny]) || (dn && mask[q + ny]) || (lf && mask[q - 1]) || (rt && mask[q + 1]);
    if (corners)
        v = v || (up && lf && mask[q - ny - 1]) || (up && rt && mask[q - ny + 1]) ||
            (dn && lf && mask[q + ny - 1]) || (dn && rt && mask[q + ny + 1]);
    out[q] = v ? 1 : 0;
}

// ---- order statistics: sorted(data)[k] and sorted(data)[k+1] by MSB-first radix select ----------
struct SelectState {
    unsigned long long prefix, mask;  // key bits fixed so far
    unsigned long long k_rem;         // rank of the wanted element among the keys matching the prefix
    unsigned long long count_le, min_gt;
    unsigned int hist[256];
};

// monotone map double -> uint64 (total order of finite values; -0.0 < +0.0)
__device__ __forceinline__ unsigned long long to_key(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_key(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void select_init_kernel(SelectState *st, long long k) {
    if (threadIdx.x == 0) {
        st->prefix = 0;
        st->mask = 0;
        st->k_rem = (unsigned long long)k;
        st->count_le = 0;
        st->min_gt = ~0ull;
    }
    st->hist[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(kTB)
select_hist_kernel(const double *__restrict__ data, long long n, int shift, SelectState *st) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix, mask = st->mask;
    for (long long q = (long long)blockIdx.x * kTB + threadIdx.x; q < n; q += (long long)gridDim.x * kTB) {
        const unsigned long long key = to_key(__ldg(data + q));
        if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void select_pick_kernel(SelectState *st, int shift) {
    // 256 threads; thread 0 walks the histogram (256 entries, eight times per call: negligible)
    __shared__ unsigned int h[256];
    h[threadIdx.x] = st->hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long k = st->k_rem;
        int b = 0;
        for (; b < 255; ++b) {
            if (k < h[b]) break;
            k -= h[b];
        }
        st->k_rem = k;
        st->prefix |= (unsigned long long)b << shift;
        st->mask |= 255ull << shift;
    }
    __syncthreads();
    st->hist[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(kTB)
select_next_kernel(const double *__restrict__ data, long long n, SelectState *st) {
    const unsigned long long kv = st->prefix;
    unsigned long long cnt = 0, mn = ~0ull;
    for (long long q = (long long)blockIdx.x * kTB + threadIdx.x; q < n; q += (long long)gridDim.x * kTB) {
        const unsigned long long key = to_key(__ldg(data + q));
        if (key <= kv) ++cnt;
        else if (key < mn) mn = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        const unsigned long long t = __shfl_down_sync(0xffffffffu, mn, o);
        mn = t < mn ? t : mn;
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(&st->count_le, cnt);
        if (mn != ~0ull) atomicMin(&st->min_gt, mn);
    }
}

__global__ void select_finish_kernel(const SelectState *st, long long k, double *out2) {
    const double v = from_key(st->prefix);
    out2[0] = v;
    // sorted[k+1]: another copy of v if more than k+1 elements are <= v, else the next larger value
    out2[1] = (st->count_le > (unsigned long long)k + 1ull || st->min_gt == ~0ull) ? v : from_key(st->min_gt);
}

inline unsigned blocks_for(long long n) { return (unsigned)((n + kTB - 1) / kTB); }

}  // namespace

void launch_c_tensor(const double *fm_aux, long long nx, long long ny, int n_aux, double h, const uint8_t *mask,
                     double *C, cudaStream_t s) {
    B2_REQUIRE((reinterpret_cast<uintptr_t>(fm_aux) & 15) == 0, "flowmap_aux must be 16-byte aligned");
    B2_REQUIRE(nx * ny < 2147483647LL * kTB, "grid too large");
    c_tensor_kernel<<<blocks_for(nx * ny), kTB, 0, s>>>(fm_aux, nx, ny, n_aux, make_divisor(2 * h), mask, C);
    B2_CHECK_CUDA(cudaGetLastError());
}

void launch_c_eig(const double *fm, long long nx, long long ny, int n_aux, double h, double dx, double dy,
                  bool aux_vecs, bool main_vals, const uint8_t *mask, double *eigvals, double *eigvecs,
                  cudaStream_t s, double *ftle, double T) {
    B2_REQUIRE((reinterpret_cast<uintptr_t>(fm) & 15) == 0 && (reinterpret_cast<uintptr_t>(eigvals) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(eigvecs) & 15) == 0,
               "flow map and eigen outputs must be 16-byte aligned");
    const int lo = (aux_vecs && main_vals) ? 2 : 1;
    c_eig_kernel<<<blocks_for(nx * ny), kTB, 0, s>>>(fm, nx, ny, n_aux, make_divisor(h != 0.0 ? 2 * h : 1.0),
                                                     make_divisor(2 * dx), make_divisor(2 * dy), aux_vecs ? 1 : 0,
                                                     main_vals ? 1 : 0, lo, mask, eigvals, eigvecs, ftle,
                                                     2 * fabs(T));
    B2_CHECK_CUDA(cudaGetLastError());
}

void launch_ftle_from_eig(const double *eigval_max, long long n, long long stride, double T, double *ftle,
                          cudaStream_t s) {
    ftle_from_eig_kernel<<<blocks_for(n), kTB, 0, s>>>(eigval_max, n, stride, 2 * fabs(T), ftle);
    B2_CHECK_CUDA(cudaGetLastError());
}

static RidgeArgs make_ridge_args(const double *f, const double *ev, long long ev_pixel_stride,
                                 long long ev_comp_stride, long long nx, long long ny, const double *x,
                                 const double *y, double dx, double dy, double sdd_thresh, double f_min) {
    RidgeArgs R{};
    R.f = f;
    R.ev = ev;
    R.x = x;
    R.y = y;
    R.ev_ps = ev_pixel_stride;
    R.ev_cs = ev_comp_stride;
    R.nx = nx;
    R.ny = ny;
    R.two_dx = make_divisor(2 * dx);
    R.two_dy = make_divisor(2 * dy);
    R.dx2 = make_divisor(dx * dx);          // dx**2
    R.dy2 = make_divisor(dy * dy);
    R.four_dxdy = make_divisor(4 * dx * dy);
    R.half_dx = dx / 2;
    R.half_dy = dy / 2;
    R.sdd_thresh = sdd_thresh;
    R.f_min = f_min;
    return R;
}

void launch_ridge_components(const double *f, const double *ev, long long ev_pixel_stride,
                             long long ev_comp_stride, long long nx, long long ny, const double *x,
                             const double *y, double dx, double dy, double sdd_thresh, double f_min,
                             double *pts_compact, long long *roots_compact, long long capacity,
                             long long *count, cudaStream_t s) {
    const RidgeArgs R = make_ridge_args(f, ev, ev_pixel_stride, ev_comp_stride, nx, ny, x, y, dx, dy,
                                        sdd_thresh, f_min);
    const long long np = nx * ny;
    B2_REQUIRE(np < 2147483647LL, "grid too large for 32-bit component labels (%lld pixels)", np);
    const unsigned nb = blocks_for(np);
    Scratch parent(sizeof(int) * np, s), counts(sizeof(int) * nb, s), offsets(sizeof(long long) * nb, s);
    int *par = static_cast<int *>(parent.ptr);
    ridge_flag_kernel<<<nb, kTB, 0, s>>>(R, par, static_cast<int *>(counts.ptr));
    scan_counts_kernel<<<1, 1024, 0, s>>>(static_cast<const int *>(counts.ptr), nb,
                                          static_cast<long long *>(offsets.ptr), count);
    ridge_merge_kernel<<<nb, kTB, 0, s>>>(par, nx, ny);
    if (pts_compact && roots_compact && capacity > 0)
        ridge_compact_roots_kernel<<<nb, kTB, 0, s>>>(R, par, static_cast<const long long *>(offsets.ptr),
                                                      pts_compact, roots_compact, capacity);
    B2_CHECK_CUDA(cudaGetLastError());
}

void launch_ridge_pts(const double *f, const double *ev, long long ev_pixel_stride, long long ev_comp_stride,
                      long long nx, long long ny, const double *x, const double *y, double dx, double dy,
                      double sdd_thresh, double f_min, double *r_pts, double *r_vec, double *sdd,
                      double *pts_compact, long long capacity, long long *count, cudaStream_t s) {
    const RidgeArgs R = make_ridge_args(f, ev, ev_pixel_stride, ev_comp_stride, nx, ny, x, y, dx, dy,
                                        sdd_thresh, f_min);
    const long long np = nx * ny;
    const bool want_count = count != nullptr;
    if (want_count && !r_pts && !r_vec && !sdd && np > 0) {
        // compact output only: the bit-mask path (one ridge test per pixel)
        const long long nb2 = (np + kRB - 1) / kRB;
        B2_REQUIRE(nb2 < 2147483647LL, "grid too large");
        Scratch bits(sizeof(unsigned) * (size_t)(nb2 * (kRB / 32)), s), counts2(sizeof(int) * (size_t)nb2, s),
            offsets2(sizeof(long long) * (size_t)nb2, s);
        ridge_detect_bits_kernel<<<(unsigned)nb2, kTB, 0, s>>>(R, static_cast<unsigned *>(bits.ptr),
                                                               static_cast<int *>(counts2.ptr));
        B2_CHECK_CUDA(cudaGetLastError());
        scan_counts_kernel<<<1, 1024, 0, s>>>(static_cast<const int *>(counts2.ptr), nb2,
                                              static_cast<long long *>(offsets2.ptr), count);
        B2_CHECK_CUDA(cudaGetLastError());
        if (pts_compact && capacity > 0) {
            ridge_compact_bits_kernel<<<(unsigned)nb2, kTB, 0, s>>>(R, static_cast<const unsigned *>(bits.ptr),
                                                                    static_cast<const long long *>(offsets2.ptr),
                                                                    pts_compact, capacity);
            B2_CHECK_CUDA(cudaGetLastError());
        }
        return;
    }
    const unsigned nb = blocks_for(np);
    Scratch counts, offsets;
    if (want_count) {
        counts = Scratch(sizeof(int) * nb, s);
        offsets = Scratch(sizeof(long long) * nb, s);
    }
    ridge_detect_kernel<<<nb, kTB, 0, s>>>(R, r_pts, r_vec, sdd, static_cast<int *>(counts.ptr));
    B2_CHECK_CUDA(cudaGetLastError());
    if (want_count) {
        scan_counts_kernel<<<1, 1024, 0, s>>>(static_cast<const int *>(counts.ptr), nb,
                                              static_cast<long long *>(offsets.ptr), count);
        B2_CHECK_CUDA(cudaGetLastError());
        if (pts_compact && capacity > 0) {
            ridge_compact_kernel<<<nb, kTB, 0, s>>>(R, static_cast<const long long *>(offsets.ptr), pts_compact,
                                                    capacity);
            B2_CHECK_CUDA(cudaGetLastError());
        }
    }
}

void launch_composition(const double *flowmaps, const double *grid6, long long nT, double *out, cudaStream_t s,
                        long long frames) {
    B2_REQUIRE(frames >= 1 && frames <= 65535, "1 .. 65535 frames per composition launch (got %lld)", frames);
    B2_REQUIRE((reinterpret_cast<uintptr_t>(flowmaps) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "flow maps must be 16-byte aligned");
    Grid2 g{};
    for (int d = 0; d < 2; ++d) {
        g.a[d] = grid6[3 * d];
        g.b[d] = grid6[3 * d + 1];
        g.n[d] = (int)grid6[3 * d + 2];
        g.delta[d] = (g.b[d] - g.a[d]) / (double)(g.n[d] - 1);
        g.inv_delta[d] = 1.0 / g.delta[d];
    }
    composition_kernel<<<dim3(blocks_for((long long)g.n[0] * g.n[1]), (unsigned)frames), kTB, 0, s>>>(
        reinterpret_cast<const double2 *>(flowmaps), g, nT, reinterpret_cast<double2 *>(out));
    B2_CHECK_CUDA(cudaGetLastError());
}

void launch_mask_dilation(const uint8_t *mask, long long nx, long long ny, bool corners, uint8_t *out,
                          cudaStream_t s) {
    mask_dilation_kernel<<<blocks_for(nx * ny), kTB, 0, s>>>(mask, nx, ny, corners ? 1 : 0, out);
    B2_CHECK_CUDA(cudaGetLastError());
}

void launch_order_stats(const double *data, long long n, long long k, double *out2, cudaStream_t s) {
    Scratch st_buf(sizeof(SelectState), s);
    SelectState *st = static_cast<SelectState *>(st_buf.ptr);
    long long want = (n + kTB - 1) / kTB;
    const unsigned nb = (unsigned)(want < 148 * 8 ? want : 148 * 8);  // grid-stride, 8 blocks per SM
    select_init_kernel<<<1, 256, 0, s>>>(st, k);
    for (int shift = 56; shift >= 0; shift -= 8) {
        select_hist_kernel<<<nb, kTB, 0, s>>>(data, n, shift, st);
        select_pick_kernel<<<1, 256, 0, s>>>(st, shift);
    }
    select_next_kernel<<<nb, kTB, 0, s>>>(data, n, st);
    select_finish_kernel<<<1, 1, 0, s>>>(st, k, out2);
    B2_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200cs
