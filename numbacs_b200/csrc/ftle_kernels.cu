// ftle_kernels.cu -- K2: fused FTLE stencil.
//
// Replaces ftle_grid_2D (/root/reference/src/numbacs/diagnostics.py:21-65): central-difference
// flow-map gradient (utils.py:41-44), Cauchy-Green tensor C = F^T F (diagnostics.py:58-59), its
// largest eigenvalue (utils.py:185-189) and log(lambda)/(2|T|) where lambda > 1, all in one pass.
// Border ring, masked pixels and lambda <= 1 pixels are exactly 0 like the reference.
//
// HBM-bound by design: 16 B read + 8 B written per pixel.  At the measured 6.5 TB/s that is one
// pixel per SM per cycle, i.e. a budget of ~100 issued instructions and ~60 FP64 instructions per
// pixel -- the naive formulation (four IEEE divisions, libm log) is FP64-bound instead.  So:
//  * a block owns a 128-column strip and walks down kRows rows with the three live flow-map rows
//    in registers: every flow-map value is fetched from L2/HBM once per strip (plus one halo row
//    per kRows); the left / right neighbours come from L1 (the row was just loaded by this block);
//  * the divisions by 2dx, 2dy are multiplications by reciprocals (<= 1 ulp per derivative);
//  * addresses are four running pointers advanced one row per iteration with the left / right
//    neighbours at immediate offsets (round 1g: 127 -> 119 instructions per pixel, 1.43 -> 1.335 ms
//    at 16384^2 = 73.6 % of the measured HBM copy bandwidth); with that the kernel is no longer
//    issue bound but memory-latency bound (ncu: long_scoreboard on top, issue slots 62 %, 64
//    registers -> 8 blocks per SM) -- a variant that saved more instructions at 72 registers and 7
//    blocks per SM was slower (1.43 ms, profiles/r1h_ftle_16384_rejected_variant.txt);
//  * log() is a 64-entry table method (top mantissa bits -> 1/c and log c from shared memory,
//    then a degree-6 log1p on |r| < 2^-7): ~9 FP64 instructions instead of ~30, error ~1e-16.
#include <cmath>
#include <cstdlib>

#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "launch.cuh"

namespace b200cs {

namespace {

#ifndef B200CS_FTLE_LEAN_SQRT
#define B200CS_FTLE_LEAN_SQRT 0   // measured slower in the register-rolling kernel (1.51 vs 1.37 ms at 16384^2)
#endif
constexpr int kCols = 128;  // threads per block = columns per strip
constexpr int kRows = 16;   // rows walked by one block

struct LogTable {
    double2 e[64];  // (1/c_i rounded, -log(1/c_i)), c_i = 1 + (i + 0.5)/64
};
__constant__ LogTable kLogTab;

// natural log of x > 0 (normal, finite) to ~1e-16: x = 2^e * m, m in [1,2)
__device__ __forceinline__ double log_table(double x, const double2 *__restrict__ tab) {
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> 14) & 63;  // top six mantissa bits
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double2 t = tab[idx];
    const double r = fma(m, t.x, -1.0);  // |r| < 2^-7
    double p = -1.0 / 6.0;
    p = fma(p, r, 1.0 / 5.0);
    p = fma(p, r, -1.0 / 4.0);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    // (double)e without the slow I2F.F64: e + 2^31 in the low word of 2^52, minus (2^52 + 2^31)
    const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;
    return fma(ed, 0.6931471805599453, fma(p, r, t.y));
}

// sqrt(x) for x >= 0 to <= 1 ulp: MUFU.RSQ64H seed (2^-22), one coupled Newton step, one residual step
__device__ __forceinline__ double sqrt_lean(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    g = fma(fma(-g, g, x), h, g);
    if (!(x < 1.0e300)) g = sqrt(x);  // overflowed gradients, inf, NaN: the IEEE routine (out of line, rare)
    return x > 1.0e-300 ? g : 0.0;   // 0, denormals (flushed by the seed) -> 0
}

__global__ void __launch_bounds__(kCols)
ftle_kernel(const double2 *__restrict__ fm, long long nx, long long ny, double scaling, double inv2dx,
            double inv2dy, const uint8_t *__restrict__ mask, double *__restrict__ out, long long row_lo,
            long long row_hi, int lo_is_border, int hi_is_border) {
    __shared__ double2 tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = kLogTab.e[threadIdx.x];
    __syncthreads();
    const long long j = (long long)blockIdx.x * kCols + threadIdx.x;
    const long long i0 = row_lo + (long long)blockIdx.y * kRows;
    if (j >= ny || i0 >= row_hi) return;
    // blockIdx.z = frame of a time series: frames are consecutive [nx, ny] grids, one mask for all
    fm += (long long)blockIdx.z * nx * ny;
    out += (long long)blockIdx.z * (row_hi - row_lo) * ny;
    const int rows = (int)((row_hi - i0 < kRows) ? (row_hi - i0) : kRows);
    const int avail = (int)((nx - i0 < kRows + 3) ? (nx - i0) : (kRows + 3));  // rows i0 .. i0+avail-1 exist
    const int r_first_border = (lo_is_border && i0 == 0) ? 0 : -1;
    const int r_last_border = (hi_is_border && nx - 1 - i0 < kRows) ? (int)(nx - 1 - i0) : -1;
    const double2 zero = make_double2(0.0, 0.0);
    // Four running pointers advanced by one row per iteration (one IMAD.WIDE each); the left /
    // right neighbours are immediate offsets -16 / +16 bytes from the row pointer.  ncu had this
    // kernel issue-bound at 127 instructions per pixel, ~30 of them 64-bit address arithmetic
    // re-derived from (base, 32-bit offset) for every access.
    const long long stride = ny;
    const double2 *pc = fm + i0 * ny + j;                 // current row, this column
    const uint8_t *pm = mask ? mask + i0 * ny + j : nullptr;
    double *po = out + (i0 - row_lo) * ny + j;
    if (j == 0 || j == ny - 1) {                          // border columns: exactly 0, no stencil
        for (int r = 0; r < rows; ++r, po += stride) *po = 0.0;
        return;
    }
    // three live rows in registers plus two rows of read-ahead
    double2 dn = (i0 >= 1) ? __ldg(pc - stride) : zero;
    double2 mid = __ldg(pc);
    double2 up = (1 < avail) ? __ldg(pc + stride) : zero;
    double2 up2 = (2 < avail) ? __ldg(pc + 2 * stride) : zero;
    const double2 *p3 = pc + 3 * stride;                  // read-ahead row
#pragma unroll 4
    for (int r = 0; r < rows; ++r) {
        const double2 up3 = (r + 3 < avail) ? __ldg(p3) : zero;
        const double2 lf = __ldg(pc - 1), rt = __ldg(pc + 1);
        bool skip = r == r_first_border || r == r_last_border;
        if (pm != nullptr) skip |= (*pm != 0);
        const double dxdx = (up.x - dn.x) * inv2dx;
        const double dxdy = (rt.x - lf.x) * inv2dy;
        const double dydx = (up.y - dn.y) * inv2dx;
        const double dydy = (rt.y - lf.y) * inv2dy;
        const double off = fma(dxdx, dxdy, dydx * dydy);
        const double a = fma(dxdx, dxdx, dydx * dydx);
        const double d = fma(dxdy, dxdy, dydy * dydy);
        const double amd = a - d;
#if B200CS_FTLE_LEAN_SQRT
        const double disc = sqrt_lean(fma(amd, amd, 4.0 * (off * off)));
#else
        const double disc = sqrt(fma(amd, amd, 4.0 * (off * off)));
#endif
        const double max_eig = 0.5 * ((a + d) + disc);
        double val = 0.0;
        // max_eig > 1 also filters NaN; huge values (overflowed gradients) go through libm
        if (!skip && max_eig > 1.0)
            val = scaling * (max_eig < 1.0e300 ? log_table(max_eig, tab) : log(max_eig));
        *po = val;
        dn = mid;
        mid = up;
        up = up2;
        up2 = up3;
        pc += stride;
        p3 += stride;
        po += stride;
        if (pm != nullptr) pm += stride;
    }
}

// ---- the TMA-fed kernel (round 2) -------------------------------------------------------------
// ncu on the kernel above at 16384^2: DRAM bytes = algorithmic, 74 % of the measured copy bandwidth,
// `long_scoreboard` on top at 40 % occupancy, 119 instructions per pixel = 0.93 issue cycles per pixel
// and SM against a budget of ~1: latency- and issue-bound at once, its read-ahead rows living in
// registers.  Here the flow map arrives through TMA instead: the [nx, 2 ny] float64 tensor is cut
// into tiles of kTR rows x 128 columns (one cp.async.bulk.tensor per tile, 16 KB), a block walks
// down a strip of 126 output columns with a ring of kSlots tiles in shared memory, two tiles always
// in flight behind an mbarrier each, so ~100 KB per SM are outstanding without a single load
// instruction or address computation in the stencil loop.  Per pixel the loop is three 16-byte
// shared-memory loads (the row above, left and right neighbours; the row below is rolled through a
// register), ~38 FP64 instructions (the square root is a MUFU seed + one Newton step + one
// residual correction, <= 1 ulp, instead of the IEEE sequence) and one store.
// Out-of-range rows / columns are zero-filled by the TMA unit; their pixels are borders (value 0)
// or are not produced at all.
#ifndef B200CS_FTLE_TR
#define B200CS_FTLE_TR 4
#endif
#ifndef B200CS_FTLE_SLOTS
#define B200CS_FTLE_SLOTS 4
#endif
constexpr int kTR = B200CS_FTLE_TR;        // rows per tile
constexpr int kSlots = B200CS_FTLE_SLOTS;  // tiles in the ring (two being consumed, the others in flight)
constexpr int kTCols = 128;     // tile columns (double2), 126 outputs + one halo column each side
constexpr int kTOut = kTCols - 2;
constexpr int kSegRows = 512;   // output rows per block

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// FP64 constants of the pixel formula, fetched ONCE per thread from the constant bank into registers:
// an FP64 instruction cannot take an immediate or a constant-bank operand on sm_100, so literals in
// the loop body are re-materialised with UMOV pairs at every use (ncu: 9.5 % of the instructions of
// the register-rolling kernel).
struct FtleConsts {
    double c6, c5, c4, c3, c2, ln2, ebias, four, half;
};
__constant__ FtleConsts kFtleC = {-1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5, 0.6931471805599453,
                                  4503601774854144.0, 4.0, 0.5};

// FTLE value of one pixel from its four neighbours (the pixel itself is not needed).  Branch-free:
// lambda <= 1, NaN and a non-finite lambda (a flow map that is not finite) give 0; `keep` = false
// (border column / masked pixel) gives 0.  tab_s = shared-window address of the log table.
__device__ __forceinline__ double ftle_pixel(const double2 up, const double2 dn, const double2 lf, const double2 rt,
                                             double inv2dx, double inv2dy, double scaling, const FtleConsts &K,
                                             unsigned tab_s, bool keep) {
    const double dxdx = (up.x - dn.x) * inv2dx;
    const double dxdy = (rt.x - lf.x) * inv2dy;
    const double dydx = (up.y - dn.y) * inv2dx;
    const double dydy = (rt.y - lf.y) * inv2dy;
    const double off = fma(dxdx, dxdy, dydx * dydy);
    const double a = fma(dxdx, dxdx, dydx * dydx);
    const double d = fma(dxdy, dxdy, dydy * dydy);
    const double amd = a - d;
    const double x = fma(amd, amd, K.four * (off * off));
    // sqrt(x): MUFU.RSQ64H seed, one coupled Newton step, one residual step (<= 1 ulp) for every
    // normal x; zero / denormal x (flushed by the seed) -> 0
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = K.half * y;
    const double r0 = fma(-h, g, K.half);
    g = fma(g, r0, g);
    h = fma(h, r0, h);
    g = fma(fma(-g, g, x), h, g);
    if (__double2hiint(x) < 0x00200000) g = 0.0;
    const double lam = K.half * ((a + d) + g);
    // log(lam) by table: top six mantissa bits -> (1/c, log c), degree-6 log1p on |r| < 2^-7
    const int hi = __double2hiint(lam);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(lam));
    double tcx, tcy;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tcx), "=d"(tcy) : "r"(tab_s + ((hi >> 10) & 0x3f0)));
    const double r = fma(m, tcx, -1.0);
    double p = fma(K.c6, r, K.c5);
    p = fma(p, r, K.c4);
    p = fma(p, r, K.c3);
    p = fma(p, r, K.c2);
    p = fma(p, r, 1.0);
    const double ed = __hiloint2double(0x43300000, ((hi >> 20) - 1023) ^ 0x80000000) - K.ebias;
    const double val = scaling * fma(ed, K.ln2, fma(p, r, tcy));
    // lam in (1, +inf) (one unsigned compare on the high word + the exact lam > 1 test): the table value
    const bool in = (unsigned)(hi - 0x3ff00000) < 0x40000000u && lam > 1.0 && keep;
    return in ? val : 0.0;
}

template <bool MASK>
__global__ void __launch_bounds__(kTCols)
ftle_tma_kernel(const __grid_constant__ CUtensorMap fm_map, long long nx, long long ny, double scaling,
                double inv2dx, double inv2dy, const uint8_t *__restrict__ mask, double *__restrict__ out,
                long long row_lo, long long row_hi, int lo_is_border, int hi_is_border) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *ring = reinterpret_cast<double2 *>(smem_raw);           // [kSlots][kTR][kTCols]
    __shared__ double2 tab[64];
    __shared__ __align__(8) unsigned long long bars[kSlots];
    const int t = threadIdx.x;
    if (t < 64) tab[t] = kLogTab.e[t];
    const FtleConsts K = kFtleC;
    const long long s0 = row_lo + (long long)blockIdx.y * kSegRows;        // first output row of this block
    const int n_out = (int)((row_hi - s0 < kSegRows) ? (row_hi - s0) : kSegRows);
    const int nchunks = (n_out + 2 + kTR - 1) / kTR;                         // loaded rows s0-1 .. s0+n_out
    const int col0 = (int)blockIdx.x * kTOut - 1;                            // tile column 0 in the grid
    constexpr unsigned kTileBytes = kTR * kTCols * sizeof(double2);
    constexpr int kTileElems = kTR * kTCols;
    if (t == 0) {
        for (int k = 0; k < kSlots; ++k) mbar_init(&bars[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c) {
        unsigned long long *bar = &bars[c % kSlots];
        mbar_expect_tx(bar, kTileBytes);
        tma_load_2d(ring + (c % kSlots) * kTileElems, &fm_map, 2 * col0, (int)(s0 - 1) + c * kTR, bar);
    };
    if (t == 0)
        for (int c = 0; c < kSlots - 1 && c < nchunks; ++c) issue(c);

    const long long j = (long long)col0 + t;
    const bool produce = t >= 1 && t <= kTOut && j < ny;
    const bool keep_col = j != 0 && j != ny - 1;                  // border columns are exactly 0
    // rows of the domain border are exactly 0: L (1-based loaded-row index of an output row) to blank
    const int L_lo_border = (lo_is_border && s0 == 0) ? 1 : -1;
    const int L_hi_border = (hi_is_border && nx - 1 >= s0 && nx - 1 < s0 + n_out) ? (int)(nx - 1 - s0) + 1 : -1;
    double *po = out + (s0 - row_lo) * ny + j - 2 * ny;            // output row L lives at po + (L + 1) * ny
    const uint8_t *pm = MASK ? mask + s0 * ny + j - 2 * ny : nullptr;
    const unsigned tab_s = smem_u32(tab);
    const double2 *base = ring + (t >= 1 ? (t <= kTOut ? t : kTOut) : 1);   // clamped: the two edge threads only tag along
    double2 dn = make_double2(0.0, 0.0), mid = dn;                 // loaded rows L-1 and L of this column
    for (int c = 0; c < nchunks; ++c) {
        const double2 *cur = base + (c % kSlots) * kTileElems;     // tile c:   loaded rows c*kTR .. c*kTR+kTR-1
        const double2 *prv = base + ((c + kSlots - 1) % kSlots) * kTileElems;   // tile c-1
        mbar_wait(&bars[c % kSlots], (unsigned)((c / kSlots) & 1));
        const int L0 = c * kTR - 1;                                // output rows L0 .. L0+kTR-1 become computable
        if (c == 0) {
            dn = cur[0];
            mid = cur[kTCols];
        }
        const bool full = c >= 1 && L0 + kTR - 1 <= n_out && L0 > L_lo_border &&
                          (L_hi_border < 0 || L0 + kTR - 1 < L_hi_border);
        // running output / mask pointers: row L0 + q is one row stride further each iteration
        po += (c == 0 ? 0 : (long long)kTR * ny);
        if (MASK) pm += (c == 0 ? 0 : (long long)kTR * ny);
        double *prow = po;
        const uint8_t *mrow = pm;
        if (full) {
            // the common case: every row of the window is an interior output row
#pragma unroll
            for (int q = 0; q < kTR; ++q) {
                const double2 up = cur[q * kTCols];                 // loaded row L+1 = c*kTR + q
                const double2 *rowL = (q == 0) ? prv + (kTR - 1) * kTCols : cur + (q - 1) * kTCols;
                const double2 lf = rowL[-1], rt = rowL[1];
                bool keep = keep_col;
                if (MASK && produce) keep = keep && (*mrow == 0);
                const double val = ftle_pixel(up, dn, lf, rt, inv2dx, inv2dy, scaling, K, tab_s, keep);
                if (produce) *prow = val;
                prow += ny;
                if (MASK) mrow += ny;
                dn = mid;
                mid = up;
            }
        } else {
            for (int q = 0; q < kTR; ++q, prow += ny, mrow += (MASK ? ny : 0)) {
                const int L = L0 + q;
                if (L < 1 || L > n_out) continue;
                const double2 up = cur[q * kTCols];
                const double2 *rowL = (q == 0) ? prv + (kTR - 1) * kTCols : cur + (q - 1) * kTCols;
                const double2 lf = rowL[-1], rt = rowL[1];
                bool keep = keep_col && L != L_lo_border && L != L_hi_border;
                if (MASK && produce) keep = keep && (*mrow == 0);
                const double val = ftle_pixel(up, dn, lf, rt, inv2dx, inv2dy, scaling, K, tab_s, keep);
                if (produce) *prow = val;
                dn = mid;
                mid = up;
            }
        }
        __syncthreads();                                              // everybody is done with tile c - 1
        if (t == 0 && c + kSlots - 1 < nchunks) issue(c + kSlots - 1);
    }
}

// cuTensorMapEncodeTiled through the runtime (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

void init_log_table() {
    static std::mutex mu;
    static std::vector<int> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    for (int d : done)
        if (d == dev) return;
    LogTable h;
    for (int i = 0; i < 64; ++i) {
        const double c = 1.0 + (i + 0.5) / 64.0;
        const double inv = 1.0 / c;  // rounded; log is taken of the ROUNDED value so the pair is consistent
        h.e[i] = make_double2(inv, (double)(-logl((long double)inv)));
    }
    B2_CHECK_CUDA(cudaMemcpyToSymbol(kLogTab, &h, sizeof(h)));
    done.push_back(dev);
}

}  // namespace

void launch_ftle(const double *fm, long long nx, long long ny, double T, double dx, double dy,
                 const uint8_t *mask, double *out, long long row_lo, long long row_hi,
                 bool lo_is_border, bool hi_is_border, cudaStream_t s, long long frames) {
    if (row_hi <= row_lo || ny <= 0 || frames <= 0) return;
    B2_REQUIRE(frames <= 65535, "too many frames for one FTLE launch (%lld)", frames);
    B2_REQUIRE((reinterpret_cast<uintptr_t>(fm) & 15) == 0, "flow map must be 16-byte aligned");
    init_log_table();
    const double scaling = 1.0 / (2.0 * fabs(T));
    const dim3 grid((unsigned)((ny + kCols - 1) / kCols), (unsigned)((row_hi - row_lo + kRows - 1) / kRows),
                    (unsigned)frames);
    B2_REQUIRE(grid.y <= 65535u, "too many rows for one FTLE launch (%lld)", row_hi - row_lo);
    B2_REQUIRE(ny * (kRows + 4) < 2147483647LL, "ny too large for the FTLE kernel (%lld)", ny);
    // TMA path: one frame, a grid large enough to fill the machine, 16-byte row pitch (always true
    // for double2 rows); everything else takes the register-rolling kernel
    static const bool tma_off = getenv("B200CS_FTLE_NO_TMA") != nullptr;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!tma_off && enc != nullptr && frames == 1 && ny >= 2 * kTOut && row_hi - row_lo >= 64 &&
        2 * ny < (1LL << 31) && nx < (1LL << 31)) {
        CUtensorMap map;
        const cuuint64_t dims[2] = {(cuuint64_t)(2 * ny), (cuuint64_t)nx};
        const cuuint64_t strides[1] = {(cuuint64_t)(2 * ny) * sizeof(double)};
        const cuuint32_t box[2] = {2 * kTCols, kTR}, estr[2] = {1, 1};
        const CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(fm), dims, strides,
                                box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        B2_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
        constexpr size_t kSmem = (size_t)kSlots * kTR * kTCols * sizeof(double2);
        static std::once_flag once[16];
        int dev = 0;
        cudaGetDevice(&dev);
        std::call_once(once[dev & 15], [] {
            cudaFuncSetAttribute(ftle_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
            cudaFuncSetAttribute(ftle_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        });
        const dim3 tgrid((unsigned)((ny + kTOut - 1) / kTOut), (unsigned)((row_hi - row_lo + kSegRows - 1) / kSegRows), 1);
        if (mask)
            ftle_tma_kernel<true><<<tgrid, kTCols, kSmem, s>>>(map, nx, ny, scaling, 1.0 / (2.0 * dx), 1.0 / (2.0 * dy),
                                                               mask, out, row_lo, row_hi, lo_is_border ? 1 : 0,
                                                               hi_is_border ? 1 : 0);
        else
            ftle_tma_kernel<false><<<tgrid, kTCols, kSmem, s>>>(map, nx, ny, scaling, 1.0 / (2.0 * dx), 1.0 / (2.0 * dy),
                                                                nullptr, out, row_lo, row_hi, lo_is_border ? 1 : 0,
                                                                hi_is_border ? 1 : 0);
        B2_CHECK_CUDA(cudaGetLastError());
        return;
    }
    ftle_kernel<<<grid, kCols, 0, s>>>(reinterpret_cast<const double2 *>(fm), nx, ny, scaling,
                                       1.0 / (2.0 * dx), 1.0 / (2.0 * dy), mask, out, row_lo, row_hi,
                                       lo_is_border ? 1 : 0, hi_is_border ? 1 : 0);
    B2_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200cs
