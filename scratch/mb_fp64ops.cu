// Microbenchmark: issue cost of DADD / DMUL / DFMA (and a butterfly-like mix) per warp instruction on one scheduler,
// for 1..4 warps per scheduler.  Prints cycles per warp-instruction per scheduler (2.0 = the nominal 16 lanes/clk).
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double *sink, long long *cyc, int iters, double seed) {
    double a[16];
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = seed + threadIdx.x * 1e-3 + j;
    const double c0 = seed * 0.999, c1 = seed * 1.001;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
            if (MODE == 0) {          // DADD, 2 register operands
#pragma unroll
                for (int j = 0; j < 16; j++) a[j] = a[j] + a[(j + 5) & 15];
            } else if (MODE == 1) {   // DMUL
#pragma unroll
                for (int j = 0; j < 16; j++) a[j] = a[j] * c0;
            } else if (MODE == 2) {   // DFMA, 3 distinct register operands
#pragma unroll
                for (int j = 0; j < 16; j++) a[j] = fma(a[(j + 3) & 15], a[(j + 7) & 15], a[j]);
            } else if (MODE == 3) {   // DFMA with one operand shared by consecutive instructions (reuse cache)
#pragma unroll
                for (int j = 0; j < 16; j++) a[j] = fma(a[j], c0, c1);
            } else if (MODE == 4) {   // radix-4 butterflies: 16 DADD on 8 values, in place
#pragma unroll
                for (int j = 0; j < 16; j += 8) {
                    double t0 = a[j] + a[j + 4], t1 = a[j] - a[j + 4], t2 = a[j + 2] + a[j + 6], t3 = a[j + 2] - a[j + 6];
                    double u0 = a[j + 1] + a[j + 5], u1 = a[j + 1] - a[j + 5], u2 = a[j + 3] + a[j + 7], u3 = a[j + 3] - a[j + 7];
                    a[j] = t0 + t2; a[j + 4] = t0 - t2; a[j + 2] = t1 - u3; a[j + 6] = t1 + u3;
                    a[j + 1] = u0 + u2; a[j + 5] = u0 - u2; a[j + 3] = u1 + t3; a[j + 7] = u1 - t3;
                }
            } else if (MODE == 5) {   // complex multiply by a constant: 2 DMUL + 2 DFMA per pair
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const double x = a[j], y = a[j + 1];
                    a[j] = x * c0 - y * c1; a[j + 1] = x * c1 + y * c0;
                }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += a[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE> void run(const char *name, int per_iter, double *sink, long long *cyc) {
    printf("%-34s", name);
    for (int threads : {128, 256, 384, 512}) {
        const int iters = 4000;
        k<MODE><<<148, threads>>>(sink, cyc, iters, 1.0000001);
        CK(cudaDeviceSynchronize());
        long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("  %dw: %5.2f", threads / 128, (double)h / iters / per_iter / (threads / 128));
    }
    printf("   cycles per warp-instruction per scheduler\n");
}
int main() {
    double *sink; long long *cyc;
    CK(cudaMalloc(&sink, 148 * 512 * 8)); CK(cudaMalloc(&cyc, 8));
    run<0>("DADD r,r", 64, sink, cyc);
    run<1>("DMUL r,c", 64, sink, cyc);
    run<2>("DFMA r,r,r (3 distinct)", 64, sink, cyc);
    run<3>("DFMA r,c,c", 64, sink, cyc);
    run<4>("radix-4 butterflies (DADD)", 64, sink, cyc);
    run<5>("complex * constant (2 DMUL+2 DFMA)", 64, sink, cyc);
    return 0;
}
