// flowmap_bickley.cu -- instantiates the flow-map kernels for one flow kind (see flowmap_kernel.cuh).
#include "flowmap_kernel.cuh"

namespace b200cs {

void launch_flowmap_bickley(const IntegArgs &A, int mode, cudaStream_t s) { launch_rhs<BickleyJet>(A, mode, s); }

void launch_lavd_bickley(const IntegArgs &A, cudaStream_t s) { launch_lavd_one<BickleyJet>(A, s); }

}  // namespace b200cs
