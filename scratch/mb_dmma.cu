// Microbenchmark: do DFMA (vector FP64 pipe) and DMMA (mma.sync m8n8k4 f64) share one pipe on B200?
// Runs DFMA-only, DMMA-only, and both interleaved; prints FMA/clk/SM for each.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>  // 0: dfma only, 1: dmma only, 2: both in every warp, 3: even warps dfma / odd warps dmma
__global__ void __launch_bounds__(256) k(double *out, double seed) {
    double a[8], b = seed, c = seed * 0.5;
    double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
    double ma = seed, mb = seed * 0.25;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * i;
    const int w = threadIdx.x >> 5;
    const bool do_f = MODE == 0 || MODE == 2 || (MODE == 3 && (w & 1) == 0);
    const bool do_m = MODE == 1 || MODE == 2 || (MODE == 3 && (w & 1) == 1);
    for (int it = 0; it < ITERS; it++) {
        if (do_f) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
        }
        if (do_m) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(d0[2 * i]), "+d"(d0[2 * i + 1]) : "d"(ma), "d"(mb));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(d1[2 * i]), "+d"(d1[2 * i + 1]) : "d"(ma), "d"(mb));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    for (int i = 0; i < 4; i++) s += d0[i] + d1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double *out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 2;
    k<MODE><<<blocks, 256>>>(out, 1e-9);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per thread per iter: dfma 8 FMA; dmma 4 instr x 256 FMA / 32 lanes = 32 FMA
    double warps = blocks * 8.0;
    double f_warps = MODE == 3 ? warps / 2 : (MODE == 1 ? 0 : warps);
    double m_warps = MODE == 3 ? warps / 2 : (MODE == 0 ? 0 : warps);
    double fma_f = f_warps * 32 * 8.0 * ITERS, fma_m = m_warps * 4 * 256.0 * ITERS;
    double clk = ms * 1e-3 * 1.965e9 * 148;
    printf("%-28s %8.3f ms  dfma %6.1f FMA/clk/SM  dmma %6.1f FMA/clk/SM  total %6.1f  (%.1f TFLOP/s)\n", name, ms, fma_f / clk, fma_m / clk,
           (fma_f + fma_m) / clk, 2 * (fma_f + fma_m) / (ms * 1e-3) * 1e-12);
}

int main() {
    double *out;
    cudaMalloc(&out, 148 * 2 * 256 * sizeof(double));
    run<0>("dfma only", out);
    run<1>("dmma only", out);
    run<2>("both, same warp", out);
    run<3>("even warps dfma, odd dmma", out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
