// flowmap_dispatch.cu -- flow-kind dispatch of the flow-map kernels + the RHS evaluation kernel.
#include "common.cuh"
#include "flows.cuh"
#include "launch.cuh"

namespace b200cs {

void launch_flowmap_dg(const IntegArgs &A, int mode, cudaStream_t s);
void launch_flowmap_bickley(const IntegArgs &A, int mode, cudaStream_t s);
void launch_flowmap_abc(const IntegArgs &A, int mode, cudaStream_t s);
void launch_flowmap_spline(int spherical, const IntegArgs &A, int mode, cudaStream_t s);
void launch_flowmap_linear(int spherical, const IntegArgs &A, int mode, cudaStream_t s);
void launch_lavd_linear(int spherical, const IntegArgs &A, cudaStream_t s);
void launch_lavd_dg(const IntegArgs &A, cudaStream_t s);
void launch_lavd_bickley(const IntegArgs &A, cudaStream_t s);
void launch_lavd_spline(int spherical, const IntegArgs &A, cudaStream_t s);

void launch_lavd_flowmap(const FlowSpec &f, const IntegArgs &A, cudaStream_t s) {
    switch (f.kind) {
    case B200CS_FLOW_DOUBLE_GYRE: launch_lavd_dg(A, s); break;
    case B200CS_FLOW_BICKLEY_JET: launch_lavd_bickley(A, s); break;
    case B200CS_FLOW_SPLINE2D: launch_lavd_spline(f.spherical, A, s); break;
    case B200CS_FLOW_LINEAR2D: launch_lavd_linear(f.spherical, A, s); break;
    default:
        set_error("LAVD needs a 2-D flow (kind %d)", f.kind);
        throw Fail{B200CS_E_INVALID};
    }
}

void launch_flowmap(const FlowSpec &f, const IntegArgs &A, int mode, cudaStream_t s) {
    switch (f.kind) {
    case B200CS_FLOW_DOUBLE_GYRE: launch_flowmap_dg(A, mode, s); break;
    case B200CS_FLOW_BICKLEY_JET: launch_flowmap_bickley(A, mode, s); break;
    case B200CS_FLOW_ABC:
        launch_flowmap_abc(A, mode, s);
        break;
    case B200CS_FLOW_SPLINE2D:
        launch_flowmap_spline(f.spherical, A, mode, s);
        break;
    case B200CS_FLOW_LINEAR2D:
        launch_flowmap_linear(f.spherical, A, mode, s);
        break;
    default:
        set_error("handle is not a flow (kind %d)", f.kind);
        throw Fail{B200CS_E_HANDLE};
    }
}

// ---------------------------------------------------------------- RHS evaluation (diagnostic)
namespace {
template <class Rhs>
__global__ void rhs_kernel(const __grid_constant__ RhsParams P, const double *__restrict__ t,
                           const double *__restrict__ y, long long npts, double *__restrict__ dy) {
    constexpr int N = Rhs::N;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npts) return;
    const Rhs rhs(P);
    double yy[N], d[N], tt[1] = {t[q]}, aux[1] = {0.0};
#pragma unroll
    for (int i = 0; i < N; ++i) yy[i] = y[q * N + i];
    if constexpr (Rhs::kAux != 0) rhs.template time_part<1>(tt, aux);
    rhs.eval(aux[0], tt[0], yy, d);
#pragma unroll
    for (int i = 0; i < N; ++i) dy[q * N + i] = d[i];
}
}  // namespace

void launch_rhs_eval(const FlowSpec &f, const RhsParams &P, const double *t, const double *y,
                     long long npts, double *dy, cudaStream_t s) {
    const int b = 128;
    const unsigned g = (unsigned)((npts + b - 1) / b);
    if (!g) return;
    switch (f.kind) {
    case B200CS_FLOW_DOUBLE_GYRE:
        if (P.p[3] != 0.0) rhs_kernel<DoubleGyreDamped><<<g, b, 0, s>>>(P, t, y, npts, dy);
        else rhs_kernel<DoubleGyre><<<g, b, 0, s>>>(P, t, y, npts, dy);
        break;
    case B200CS_FLOW_BICKLEY_JET: rhs_kernel<BickleyJet><<<g, b, 0, s>>>(P, t, y, npts, dy); break;
    case B200CS_FLOW_ABC: rhs_kernel<Abc><<<g, b, 0, s>>>(P, t, y, npts, dy); break;
    case B200CS_FLOW_SPLINE2D:
        if (f.spherical == 1) rhs_kernel<Spline2D<1>><<<g, b, 0, s>>>(P, t, y, npts, dy);
        else if (f.spherical == 2) rhs_kernel<Spline2D<2>><<<g, b, 0, s>>>(P, t, y, npts, dy);
        else rhs_kernel<Spline2D<0>><<<g, b, 0, s>>>(P, t, y, npts, dy);
        break;
    case B200CS_FLOW_LINEAR2D:
        if (f.spherical == 1) rhs_kernel<Spline2D<1, true>><<<g, b, 0, s>>>(P, t, y, npts, dy);
        else if (f.spherical == 2) rhs_kernel<Spline2D<2, true>><<<g, b, 0, s>>>(P, t, y, npts, dy);
        else rhs_kernel<Spline2D<0, true>><<<g, b, 0, s>>>(P, t, y, npts, dy);
        break;
    default:
        set_error("handle is not a flow (kind %d)", f.kind);
        throw Fail{B200CS_E_HANDLE};
    }
    B2_CHECK_CUDA(cudaGetLastError());
}

}  // namespace b200cs
