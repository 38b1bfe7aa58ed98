// flowmap_spline.cu -- instantiates the flow-map kernels for one flow kind (see flowmap_kernel.cuh).
#include "flowmap_kernel.cuh"

namespace b200cs {

void launch_flowmap_spline(int spherical, const IntegArgs &A, int mode, cudaStream_t s) {
    if (spherical == 1) launch_rhs<Spline2D<1>>(A, mode, s);
    else if (spherical == 2) launch_rhs<Spline2D<2>>(A, mode, s);
    else launch_rhs<Spline2D<0>>(A, mode, s);
}

void launch_lavd_spline(int spherical, const IntegArgs &A, cudaStream_t s) {
    if (spherical == 1) launch_lavd_one<Spline2D<1>>(A, s);
    else if (spherical == 2) launch_lavd_one<Spline2D<2>>(A, s);
    else launch_lavd_one<Spline2D<0>>(A, s);
}

}  // namespace b200cs
