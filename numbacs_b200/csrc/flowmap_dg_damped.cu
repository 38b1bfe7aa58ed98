// flowmap_dg_damped.cu -- the double gyre with alpha != 0 (flows.py:1157-1158 with the
// -alpha*y terms); see flowmap_dg.cu.
#include "flowmap_kernel.cuh"

namespace b200cs {

void launch_flowmap_dg_damped(const IntegArgs &A, int mode, cudaStream_t s) {
    launch_rhs<DoubleGyreDamped>(A, mode, s);
}

void launch_lavd_dg_damped(const IntegArgs &A, cudaStream_t s) { launch_lavd_one<DoubleGyreDamped>(A, s); }

}  // namespace b200cs
