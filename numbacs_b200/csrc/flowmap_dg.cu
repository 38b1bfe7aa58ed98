// flowmap_dg.cu -- instantiates the flow-map kernels for one flow kind (see flowmap_kernel.cuh).
#include "flowmap_kernel.cuh"

namespace b200cs {

void launch_flowmap_dg(const IntegArgs &A, int mode, cudaStream_t s) { launch_rhs<DoubleGyre>(A, mode, s); }

void launch_lavd_dg(const IntegArgs &A, cudaStream_t s) { launch_lavd_one<DoubleGyre>(A, s); }

}  // namespace b200cs
