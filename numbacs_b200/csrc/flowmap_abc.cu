// flowmap_abc.cu -- instantiates the flow-map kernels for one flow kind (see flowmap_kernel.cuh).
#include "flowmap_kernel.cuh"

namespace b200cs {

void launch_flowmap_abc(const IntegArgs &A, int mode, cudaStream_t s) {
    B2_REQUIRE(!mode, "abc is a 3-D flow: use the point-list entry (flowmap / flowmap_n)");
    launch_rhs<Abc>(A, false, s);
}

}  // namespace b200cs
