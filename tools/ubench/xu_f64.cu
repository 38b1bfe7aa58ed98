// Which pipe do FP64 conversions / roundings use on sm_100a?  (scratch micro-benchmark)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 xu_f64.cu -o xu_f64
// Every thread runs 4 independent DFMA chains (enough to saturate the FP64 pipe at 5 warps per
// sub-partition) and, per 8 DFMAs, EXTRA instructions of one kind.  If the extra instruction runs on
// another pipe (XU) the time per iteration does not move; if it shares the FP64 pipe it grows by
// EXTRA/8 (or more, for a multi-pass instruction).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int EXTRA>
__global__ void __launch_bounds__(128) k(int iters, double *out, double seed) {
    double a[4];
    for (int i = 0; i < 4; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
    const double m = 0.9999999, b = 1e-7;
    double e = seed * 3.7 + threadIdx.x;
    int ei = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = fma(a[i], m, b);
#pragma unroll
        for (int x = 0; x < EXTRA; ++x) {
            if (MODE == 1) { e = rint(e); e = __hiloint2double(__double2hiint(e) ^ 0x100, __double2loint(e)); }  // FRND.F64 + LOP3
            if (MODE == 2) { ei += __double2int_rn(e); e = __hiloint2double(__double2hiint(e) ^ (ei & 1), __double2loint(e)); }  // F2I.F64
            if (MODE == 3) { e = (double)ei; ei = __double2loint(e) + it; }   // I2F.F64.S32
            if (MODE == 4) { e = a[0] + e; }                         // DADD (same pipe, control)
            if (MODE == 5) { float f = __double2float_rn(e); e = __hiloint2double(__double2hiint(e), __float_as_int(f)); }  // F2F.F32.F64
            if (MODE == 6) { asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(e) : "d"(e)); }  // MUFU.RCP64H
        }
    }
    double s = e + ei;
    for (int i = 0; i < 4; ++i) s += a[i];
    if (s == 123.456) *out = s;
}

template <int MODE, int EXTRA>
float run(int iters, double *o) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, EXTRA><<<148 * 5, 128>>>(iters, o, 1.0);
    cudaEventRecord(e0);
    k<MODE, EXTRA><<<148 * 5, 128>>>(iters, o, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    double *o;
    cudaMalloc(&o, 8);
    const int iters = 200000;
    const float base = run<0, 0>(iters, o);
    printf("base (8 DFMA/iter, 5 warps/SMSP): %.3f ms  -> %.2f cycles per DFMA per SMSP at 1.965 GHz\n", base,
           base * 1e-3 * 1.965e9 / (iters * 8.0 * 5));
    const char *names[] = {"", "FRND.F64", "F2I.F64", "I2F.F64", "DADD", "F2F.F32.F64", "MUFU.RCP64H"};
#define ROW(M)                                                                                           \
    printf("%-12s extra=1: %.3f  extra=2: %.3f  extra=4: %.3f   (x base)\n", names[M], run<M, 1>(iters, o) / base, \
           run<M, 2>(iters, o) / base, run<M, 4>(iters, o) / base);
    ROW(1) ROW(2) ROW(3) ROW(4) ROW(5) ROW(6)
    return 0;
}
