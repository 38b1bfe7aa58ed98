// ftle_kernels.cu -- K2: fused FTLE stencil.
//
// Replaces ftle_grid_2D (/root/reference/src/numbacs/diagnostics.py:21-65): central-difference
// flow-map gradient (utils.py:41-44), Cauchy-Green tensor C = F^T F (diagnostics.py:58-59), its
// largest eigenvalue (utils.py:185-189) and log(lambda)/(2|T|) where lambda > 1, all in one pass.
// Border ring, masked pixels and lambda <= 1 pixels are exactly 0 like the reference.
//
// The kernel is HBM-bound by design (16 B read + 8 B written per pixel): a (16+2) x (64+2) tile
// of double2 flow-map values is staged in shared memory with coalesced 16-byte loads, so every
// value is fetched from L2/HBM once per tile (halo re-reads hit L2).  The two divisions by 2dx,
// 2dy become multiplications by their reciprocals: a <= 1 ulp change per derivative that keeps
// the FP64 pipe (sqrt + log + ~30 flops per pixel) below the memory time.
#include "common.cuh"
#include "launch.cuh"

namespace b200cs {

namespace {

constexpr int kTJ = 64;   // tile columns (j, contiguous)
constexpr int kTI = 16;   // tile rows (i)
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
ftle_kernel(const double2 *__restrict__ fm, long long nx, long long ny, double scaling, double inv2dx,
            double inv2dy, const uint8_t *__restrict__ mask, double *__restrict__ out, long long row_lo,
            long long row_hi, int lo_is_border, int hi_is_border) {
    __shared__ double2 tile[kTI + 2][kTJ + 2];
    const long long i0 = row_lo + (long long)blockIdx.y * kTI;
    const long long j0 = (long long)blockIdx.x * kTJ;
    // stage the tile plus a one-cell halo; out-of-slab cells are never used
    for (int idx = threadIdx.x; idx < (kTI + 2) * (kTJ + 2); idx += kThreads) {
        const int li = idx / (kTJ + 2), lj = idx - li * (kTJ + 2);
        const long long gi = i0 - 1 + li, gj = j0 - 1 + lj;
        double2 v = make_double2(0.0, 0.0);
        if (gi >= 0 && gi < nx && gj >= 0 && gj < ny) v = __ldg(fm + gi * ny + gj);
        tile[li][lj] = v;
    }
    __syncthreads();
    const int tj = threadIdx.x % kTJ;
    const int ti0 = threadIdx.x / kTJ;  // 0..3
    const long long j = j0 + tj;
    if (j >= ny) return;
#pragma unroll
    for (int ti = ti0; ti < kTI; ti += kThreads / kTJ) {
        const long long i = i0 + ti;
        if (i >= row_hi) break;
        double val = 0.0;
        const bool border = (j == 0) || (j == ny - 1) || (i == 0 && lo_is_border) ||
                            (i == nx - 1 && hi_is_border);
        if (!border && !(mask != nullptr && mask[i * ny + j])) {
            const double2 up = tile[ti + 2][tj + 1], dn = tile[ti][tj + 1];
            const double2 rt = tile[ti + 1][tj + 2], lf = tile[ti + 1][tj];
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
            if (max_eig > 1.0) val = scaling * log(max_eig);
        }
        out[(i - row_lo) * ny + j] = val;
    }
}

}  // namespace

void launch_ftle(const double *fm, long long nx, long long ny, double T, double dx, double dy,
                 const uint8_t *mask, double *out, long long row_lo, long long row_hi,
                 bool lo_is_border, bool hi_is_border, cudaStream_t s) {
    if (row_hi <= row_lo || ny <= 0) return;
    B2_REQUIRE((reinterpret_cast<uintptr_t>(fm) & 15) == 0, "flow map must be 16-byte aligned");
    const double scaling = 1.0 / (2.0 * fabs(T));
    const dim3 grid((unsigned)((ny + kTJ - 1) / kTJ), (unsigned)((row_hi - row_lo + kTI - 1) / kTI));
    B2_REQUIRE(grid.y <= 65535u, "too many rows for one FTLE launch (%lld)", row_hi - row_lo);
    ftle_kernel<<<grid, kThreads, 0, s>>>(reinterpret_cast<const double2 *>(fm), nx, ny, scaling,
                                          1.0 / (2.0 * dx), 1.0 / (2.0 * dy), mask, out, row_lo, row_hi,
                                          lo_is_border ? 1 : 0, hi_is_border ? 1 : 0);
    B2_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200cs
