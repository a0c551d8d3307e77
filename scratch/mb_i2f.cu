// Does I2F.F64.S32 share the FP64 pipe with DFMA on B200?  One conversion per 14 DFMAs (the ratio of the forward FFT of the PBS
// kernel: 32 digit conversions per 471 FP64 instructions), conversion done (0) with the 2^52 trick + one DADD, (1) with I2F.F64.S32,
// (2) not at all (DFMAs only).  8 warps per SM sub-partition... see the printed table.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(512) k(double *out, int seed) {
    double d[14]; int a = seed + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 14; i++) d[i] = seed * 1.37 + i;
    for (int it = 0; it < ITERS; it++) {
        double t;
        if (MODE == 0) t = __hiloint2double(0x43300000, a & 255) - 4503599627370624.0;
        else if (MODE == 1) t = (double)((a & 255) - 128);
        else t = 1e-9;
        a = a * 1664525 + 1013904223;
        d[0] = fma(d[0], 1.0000001, t);
#pragma unroll
        for (int i = 1; i < 14; i++) d[i] = fma(d[i], 1.0000001, 1e-9);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 14; i++) s += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + a;
}
template <int MODE>
void run(const char *name, double *out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148;
    k<MODE><<<blocks, 512>>>(out, 3);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 512>>>(out, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %8.3f ms  %7.2f cycles per iteration per scheduler (4 warps each)\n", name, ms, ms * 1e-3 * 1.965e9 / ITERS);
}
int main() {
    double *out; cudaMalloc(&out, 148 * 512 * 8);
    run<2>("14 DFMA", out);
    run<0>("14 DFMA + 2^52 trick (1 DADD)", out);
    run<1>("14 DFMA + IADD + I2F.F64.S32", out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
