// microbenchmarks: legacy mma.sync int8 rate, IMAD / IMAD.WIDE rates on sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) imma_kernel(int *out, int iters) {
    int c[8][4];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    if (s == 0x12345) out[0] = s;
}
__global__ void __launch_bounds__(256) imad_kernel(unsigned long long *out, int iters, unsigned u, unsigned long long k) {
    unsigned long long acc[8];
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    unsigned klo = (unsigned)k, khi = (unsigned)(k >> 32);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {   // one u64 MAC: wide lo + 32-bit hi
            unsigned lo = (unsigned)acc[i], hi = (unsigned)(acc[i] >> 32);
            unsigned long long t = (unsigned long long)u * klo + acc[i];
            unsigned h2 = u * khi + (unsigned)(t >> 32);
            acc[i] = ((unsigned long long)h2 << 32) | (unsigned)t;
            (void)lo; (void)hi;
        }
    }
    unsigned long long s = 0;
    for (int i = 0; i < 8; i++) s += acc[i];
    if (s == 0x12345) out[0] = s;
}
__global__ void __launch_bounds__(256) imad32_kernel(unsigned *out, int iters, unsigned u, unsigned k) {
    unsigned acc[16];
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = acc[i] * u + k;
    }
    unsigned s = 0;
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 0x12345) out[0] = s;
}
__global__ void __launch_bounds__(256) imadwide_kernel(unsigned long long *out, int iters, unsigned u, unsigned k) {
    unsigned long long acc[16];
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = (unsigned long long)((unsigned)acc[i] ^ u) * k + acc[i];
    }
    unsigned long long s = 0;
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 0x12345) out[0] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void *d; cudaMalloc(&d, 64);
    const int iters = 20000, grid = sms * 4;
    float ms = timeit([&] { imma_kernel<<<grid, 256>>>((int *)d, iters); });
    double macs = (double)grid * 8 /*warps*/ * iters * 8 * 4096.0;
    printf("imma m16n8k32 s8: %.3f ms, %.1f int8 TMAC/s, %.0f MAC/clk/SM @1.965GHz\n", ms, macs / ms * 1e-9, macs / (ms * 1e-3) / sms / 1.965e9);
    ms = timeit([&] { imad_kernel<<<grid, 256>>>((unsigned long long *)d, iters, 3001, 0x123456789abcdefull); });
    double m64 = (double)grid * 256 * iters * 8;
    printf("u64 MAC (IMAD.WIDE+IMAD): %.3f ms, %.2f GMAC/s, %.1f MAC/clk/SM\n", ms, m64 / ms * 1e-6, m64 / (ms * 1e-3) / sms / 1.965e9);
    ms = timeit([&] { imad32_kernel<<<grid, 256>>>((unsigned *)d, iters, 3001, 77); });
    double m32 = (double)grid * 256 * iters * 16;
    printf("IMAD 32: %.3f ms, %.1f op/clk/SM\n", ms, m32 / (ms * 1e-3) / sms / 1.965e9);
    ms = timeit([&] { imadwide_kernel<<<grid, 256>>>((unsigned long long *)d, iters, 3001, 77); });
    printf("IMAD.WIDE(+LOP): %.3f ms, %.1f op/clk/SM\n", ms, m32 / (ms * 1e-3) / sms / 1.965e9);
    return 0;
}
