// flowmap_kernel.cuh -- K1 / K1n: the particle flow-map kernel template (one thread per particle).
//
// Replaces the prange loops of /root/reference/src/numbacs/integration.py:
//   flowmap 46-52, flowmap_n 105-111, flowmap_grid_2D 166-172, flowmap_n_grid_2D 517-523.
// Thread q owns particle q of the C-order 'ij' grid (j fastest), so a warp is 32 consecutive
// y-neighbours: their trajectories, step counts and (for the spline flow) coefficient cells are
// similar, and the final 16-byte stores are fully coalesced.
// One translation unit per flow kind instantiates it (flowmap_<kind>.cu) so they build in parallel.
#pragma once
#include "common.cuh"
#include "dop853.cuh"
#include "flows.cuh"
#include "launch.cuh"

namespace b200cs {

namespace {

// Launch shape per (flow, output mode).  Default: 128-thread blocks, as many as the registers allow.
// The final-time double-gyre kernel (the headline workload) runs ONE 640-thread block per SM
// (20 warps at <= 96 registers) with a block barrier before every step attempt: all 20 warps walk
// the 62 KB unrolled body together, which removes most of its instruction-cache misses.
template <class Rhs, bool DENSE>
struct KernelShape {
    static constexpr int kThreads = 128;
    static constexpr int kMinBlocks = 1;
    static constexpr bool kLockstep = false;
};
template <>
struct KernelShape<DoubleGyre, false> {
    static constexpr int kThreads = 640;
    static constexpr int kMinBlocks = 1;
    static constexpr bool kLockstep = true;
};

template <int N>
struct RowSink {
    double *row;  // out + q*n*N
    bool vec;     // N == 2 and the base pointer is 16-byte aligned
    __device__ __forceinline__ void operator()(int k, const double (&v)[N]) const {
        if (N == 2 && vec) {
            *reinterpret_cast<double2 *>(row + 2 * (long long)k) = make_double2(v[0], v[1]);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) row[(long long)k * N + i] = v[i];
        }
    }
};

template <class Rhs, bool DENSE, bool GRID>
__global__ void __launch_bounds__(KernelShape<Rhs, DENSE>::kThreads, KernelShape<Rhs, DENSE>::kMinBlocks)
flowmap_kernel(const __grid_constant__ IntegArgs A) {
    constexpr int N = Rhs::N;
    constexpr int kBlock = KernelShape<Rhs, DENSE>::kThreads;
    constexpr bool kLockstep = KernelShape<Rhs, DENSE>::kLockstep;
    const long long q = (long long)blockIdx.x * kBlock + threadIdx.x;
    const bool in_range = q < A.npts;
    bool active = in_range;
    if (active && A.mask != nullptr) active = (A.mask[q] == 0);

    double y[N];
#pragma unroll
    for (int i = 0; i < N; ++i) y[i] = 0.0;
    if (active) {
        if (GRID) {
            const long long i = q / A.ny, j = q - i * A.ny;
            y[0] = A.x[i];
            y[1] = A.y[j];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) y[i] = A.pts[q * N + i];
        }
    }

    StepCounts cnt;
    int status = B200CS_ST_MASKED;
    const long long row_len = DENSE ? (long long)A.n_out * N : N;
    double *row = A.out + q * row_len;

    const Rhs rhs(A.rhs);
    const bool integrate = active && (A.xend != A.x0);   // T == 0: the flow map is the identity
    if (DENSE) {
        RowSink<N> sink{row, A.out_aligned16 != 0};
        if (active) {
            sink(0, y);  // row 0 is the initial condition
            if (!integrate)
                for (int k = 1; k < A.n_out; ++k) sink(k, y);
        } else if (in_range) {
            for (long long k = 0; k < row_len; ++k) row[k] = 0.0;  // masked: zeros (integration.py:163, 515)
        }
        status = dop853_integrate<true, kLockstep>(rhs, integrate, y, A.x0, A.xend, A.rtol, A.atol, A.n_out,
                                                   A.out_p0, A.out_t0, A.out_step, sink, cnt);
    } else {
        status = dop853_integrate<false, kLockstep>(rhs, integrate, y, A.x0, A.xend, A.rtol, A.atol, 0, 0.0, 0.0,
                                                    0.0, NoSink<N>{}, cnt);
    }
    if (active && !integrate) status = B200CS_ST_OK;

    if (in_range) {
        if (!DENSE) {
            if (N == 2 && A.out_aligned16) {
                *reinterpret_cast<double2 *>(row) = make_double2(y[0], y[1]);
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) row[i] = y[i];
            }
        }
        if (A.status) A.status[q] = status;
        if (A.steps) {
            A.steps[2 * q] = cnt.accepted;
            A.steps[2 * q + 1] = cnt.rejected;
        }
    }
    if (A.stats) {
        const int nstep = cnt.accepted + cnt.rejected;
        unsigned long long nfev =
            active && (A.xend != A.x0) ? 2ull + 11ull * nstep + cnt.accepted + 3ull * cnt.dense : 0ull;
        unsigned long long acc = cnt.accepted, rej = cnt.rejected;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nfev += __shfl_down_sync(0xffffffffu, nfev, o);
            acc += __shfl_down_sync(0xffffffffu, acc, o);
            rej += __shfl_down_sync(0xffffffffu, rej, o);
        }
        if ((threadIdx.x & 31) == 0 && nfev) {
            atomicAdd(&A.stats[0], nfev);
            atomicAdd(&A.stats[1], acc);
            atomicAdd(&A.stats[2], rej);
        }
    }
}

template <class Rhs, bool DENSE, bool GRID>
void launch_one(const IntegArgs &A, cudaStream_t s) {
    constexpr int kBlock = KernelShape<Rhs, DENSE>::kThreads;
    const long long blocks = (A.npts + kBlock - 1) / kBlock;
    if (blocks <= 0) return;
    B2_REQUIRE(blocks < 2147483647LL, "too many particles for one launch (%lld)", A.npts);
    flowmap_kernel<Rhs, DENSE, GRID><<<(unsigned)blocks, kBlock, 0, s>>>(A);
    B2_CHECK_CUDA(cudaGetLastError());
}

template <class Rhs>
void launch_rhs(const IntegArgs &A, bool grid_mode, cudaStream_t s) {
    const bool dense = A.n_out >= 2;
    if (dense) {
        if (grid_mode) launch_one<Rhs, true, true>(A, s);
        else launch_one<Rhs, true, false>(A, s);
    } else {
        if (grid_mode) launch_one<Rhs, false, true>(A, s);
        else launch_one<Rhs, false, false>(A, s);
    }
}

}  // namespace

}  // namespace b200cs
