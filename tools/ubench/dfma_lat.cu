// DFMA latency / throughput micro-benchmark (scratch tool): nvcc -arch=sm_100a dfma_lat.cu -o dfma_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(int iters, double *out, long long *cyc) {
    double a[CHAINS];
    for (int i = 0; i < CHAINS; ++i) a[i] = 1.0 + i + threadIdx.x * 1e-3;
    const double m = 0.9999999, b = 1e-7;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (s == 123.0) *out = s;
}
template <int CHAINS>
void run(int warps_per_block) {
    double *o; long long *c, h;
    cudaMalloc(&o, 8); cudaMalloc(&c, 8);
    int iters = 2000;
    k<CHAINS><<<1, 32 * warps_per_block>>>(iters, o, c);
    k<CHAINS><<<1, 32 * warps_per_block>>>(iters, o, c);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / (iters * 8.0 * CHAINS);
    printf("chains %d warps/SM %d (per SMSP %.2f): %.2f cycles per DFMA per warp; SMSP DFMA issue interval %.2f cycles\n",
           CHAINS, warps_per_block, warps_per_block / 4.0, per, per / (warps_per_block / 4.0 < 1 ? 1 : warps_per_block / 4.0));
}
int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1);
    run<1>(4); run<1>(8); run<1>(16); run<1>(32);
    run<2>(4); run<2>(8); run<2>(16); run<4>(4); run<4>(8); run<4>(16);
    return 0;
}
