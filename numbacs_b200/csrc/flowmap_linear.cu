// flowmap_linear.cu -- instantiates the flow-map kernels for the trilinear gridded flow
// (get_flow_linear_2D, flows.py:418-506; see flowmap_kernel.cuh).
#include "flowmap_kernel.cuh"

namespace b200cs {

void launch_flowmap_linear(int spherical, const IntegArgs &A, int mode, cudaStream_t s) {
    if (spherical == 1) launch_rhs<Spline2D<1, true>>(A, mode, s);
    else if (spherical == 2) launch_rhs<Spline2D<2, true>>(A, mode, s);
    else launch_rhs<Spline2D<0, true>>(A, mode, s);
}

void launch_lavd_linear(int spherical, const IntegArgs &A, cudaStream_t s) {
    if (spherical == 1) launch_lavd_one<Spline2D<1, true>>(A, s);
    else if (spherical == 2) launch_lavd_one<Spline2D<2, true>>(A, s);
    else launch_lavd_one<Spline2D<0, true>>(A, s);
}

}  // namespace b200cs
