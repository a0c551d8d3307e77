// Throughput of the FP64 conversion instructions the PBS kernel uses, per SM per clock (B200).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int seed) {
    int a[8]; double d[8]; long long l[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; d[i] = seed * 1.37 + i; l[i] = 0; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { d[i] += (double)a[i]; a[i] ^= it; }                    // I2F.F64.S32 (+DADD)
            if (MODE == 1) { d[i] = rint(d[i]) * 1.0000001; }                          // FRND.F64 (+DMUL)
            if (MODE == 2) { l[i] += __double2ll_rn(d[i]); d[i] += 1.5; }              // F2I.S64.F64 (+DADD)
            if (MODE == 3) { d[i] = fma(d[i], 1.0000001, 1e-9); }                      // DFMA only (reference)
            if (MODE == 4) { d[i] += __hiloint2double(0x43300000, a[i]); a[i] ^= it; } // magic (DADD only)
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i] + a[i] + (double)l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, double *out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4;
    k<MODE><<<blocks, 256>>>(out, 3);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * 8.0 * ITERS;
    printf("%-34s %8.3f ms  %7.1f ops/clk/SM\n", name, ms, ops / (ms * 1e-3 * 1.965e9 * 148));
}
int main() {
    double *out; cudaMalloc(&out, 148 * 4 * 256 * 8);
    run<3>("DFMA", out);
    run<0>("I2F.F64.S32 + DADD", out);
    run<4>("hiloint2double + DADD", out);
    run<1>("FRND.F64 + DMUL", out);
    run<2>("F2I.S64.F64 + DADD", out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
