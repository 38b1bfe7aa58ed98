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

constexpr int kBlock = 128;

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
__global__ void __launch_bounds__(kBlock) flowmap_kernel(const __grid_constant__ IntegArgs A) {
    constexpr int N = Rhs::N;
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

    if (active) {
        const Rhs rhs(A.rhs);
        if (DENSE) {
            RowSink<N> sink{row, A.out_aligned16 != 0};
            sink(0, y);  // row 0 is the initial condition
            if (A.xend == A.x0) {
                status = B200CS_ST_OK;
                for (int k = 1; k < A.n_out; ++k) sink(k, y);
            } else {
                status = dop853_integrate<true>(rhs, y, A.x0, A.xend, A.rtol, A.atol, A.n_out, A.out_p0,
                                                A.out_t0, A.out_step, sink, cnt);
            }
        } else {
            status = (A.xend == A.x0)
                         ? B200CS_ST_OK
                         : dop853_integrate<false>(rhs, y, A.x0, A.xend, A.rtol, A.atol, 0, 0.0, 0.0, 0.0,
                                                   NoSink<N>{}, cnt);
        }
    } else if (in_range && DENSE) {
        for (long long k = 0; k < row_len; ++k) row[k] = 0.0;  // masked: zeros (integration.py:163, 515)
    }

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

template <class Rhs>
void launch_rhs(const IntegArgs &A, bool grid_mode, cudaStream_t s) {
    const long long blocks = (A.npts + kBlock - 1) / kBlock;
    if (blocks <= 0) return;
    B2_REQUIRE(blocks < 2147483647LL, "too many particles for one launch (%lld)", A.npts);
    const dim3 g((unsigned)blocks), b(kBlock);
    const bool dense = A.n_out >= 2;
    if (dense) {
        if (grid_mode) flowmap_kernel<Rhs, true, true><<<g, b, 0, s>>>(A);
        else flowmap_kernel<Rhs, true, false><<<g, b, 0, s>>>(A);
    } else {
        if (grid_mode) flowmap_kernel<Rhs, false, true><<<g, b, 0, s>>>(A);
        else flowmap_kernel<Rhs, false, false><<<g, b, 0, s>>>(A);
    }
    B2_CHECK_CUDA(cudaGetLastError());
}

}  // namespace

}  // namespace b200cs
