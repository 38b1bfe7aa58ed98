// flowmap_dg.cu -- instantiates the flow-map kernels for one flow kind (see flowmap_kernel.cuh).
// The damped double gyre (alpha != 0) is its own instantiation in flowmap_dg_damped.cu, so the
// default flow's hot loop carries no damping terms.
#include "flowmap_kernel.cuh"

namespace b200cs {

void launch_flowmap_dg_damped(const IntegArgs &A, int mode, cudaStream_t s);
void launch_lavd_dg_damped(const IntegArgs &A, cudaStream_t s);

void launch_flowmap_dg(const IntegArgs &A, int mode, cudaStream_t s) {
    if (A.rhs.p[3] != 0.0) launch_flowmap_dg_damped(A, mode, s);
    else launch_rhs<DoubleGyre>(A, mode, s);
}

void launch_lavd_dg(const IntegArgs &A, cudaStream_t s) {
    if (A.rhs.p[3] != 0.0) launch_lavd_dg_damped(A, s);
    else launch_lavd_one<DoubleGyre>(A, s);
}

}  // namespace b200cs
