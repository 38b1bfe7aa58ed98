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

#include "common.cuh"
#include "launch.cuh"

namespace b200cs {

namespace {

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
    return fma((double)e, 0.6931471805599453, fma(p, r, t.y));
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
        const double disc = sqrt(fma(amd, amd, 4.0 * (off * off)));
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
    ftle_kernel<<<grid, kCols, 0, s>>>(reinterpret_cast<const double2 *>(fm), nx, ny, scaling,
                                       1.0 / (2.0 * dx), 1.0 / (2.0 * dy), mask, out, row_lo, row_hi,
                                       lo_is_border ? 1 : 0, hi_is_border ? 1 : 0);
    B2_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200cs
