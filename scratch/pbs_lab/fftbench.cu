// Microbenchmark: throughput of the forward-FFT instruction stream of the PBS kernel (digits -> pass 1 -> exchange ->
// pass 2 -> hand-over store) as a function of the number of warps per scheduler, with no barriers and no MAC role.
//   ./fftbench
#include <cstdio>
#include <cstdlib>
#include "../../tfhe-aes_b200/csrc/cmux_core.cuh"
#include "../../tfhe-aes_b200/csrc/twiddle_host.h"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(512, 1) fftbench(const cd *tw_g, int iters, double *sink, long long *cyc) {
    extern __shared__ __align__(128) unsigned char raw[];
    cd *tw = reinterpret_cast<cd *>(raw);
    cd *xb = tw + 256;
    const int tid = threadIdx.x, lane = tid & 15, gid = tid >> 4;
    for (int i = tid; i < 256; i += blockDim.x) tw[i] = tw_g[i];
    cd *slot = xb + gid * 256;
    for (int i = lane; i < 256; i += 16) slot[i] = cmk(0.0, 0.0);
    __syncthreads();
    cd v[16];
    uint32_t st_re[16], st_im[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { st_re[k] = tid * 2654435761u + k * 40503u; st_im[k] = tid * 40503u + k * 2654435761u; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE != 2) next_digits<8, 5>(v, st_re, st_im, 1 + (it & 3));
        if (MODE == 0) {
            fft256_fwd_pass1_compute(v, lane, tw);
            fft256_fwd_pass1_store(v, lane, slot);
            __syncwarp();
            fft256_fwd_pass2(v, lane, slot);
        } else if (MODE == 1) {
            fft256_fwd_pass1_raw(v);
            fft256_fwd_pass1_store(v, lane, slot);
            __syncwarp();
            fft256_fwd_pass2_tw(v, lane, tw, slot);
        } else if (MODE == 2) {   // FP64 only: the two 16-point transforms on registers
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = cmk((double)(st_re[k] & 255), (double)(st_im[k] & 255));
            fft256_fwd_pass1_raw(v);
            fft16<1>(v);
        }
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) slot[lane + 16 * k2] = v[rev4(k2)];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++) { st_re[k] += __double2loint(v[k].x); }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s += v[k].x + v[k].y + st_re[k];
    sink[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const cd *tw, double *sink, long long *cyc, const char *name) {
    for (int threads : {128, 256, 384, 512}) {
        const int iters = 2000;
        const size_t smem = (256 + (threads / 16) * 256) * sizeof(cd);
        CK(cudaFuncSetAttribute(fftbench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fftbench<MODE><<<148, threads, smem>>>(tw, iters, sink, cyc);
        CK(cudaDeviceSynchronize());
        long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double per_it = (double)h / iters;
        // per scheduler: (threads/128) warps, each iteration = one pass of the level loop for 2 polynomials
        printf("%-10s %3d threads (%d warps/scheduler): %7.0f cycles per iteration, %6.0f cycles per warp-iteration per scheduler\n", name, threads,
               threads / 128, per_it, per_it / (threads / 128));
    }
}
int main() {
    cd htw[256]; make_twiddle_table(htw);
    cd *tw; double *sink; long long *cyc;
    CK(cudaMalloc(&tw, sizeof(htw))); CK(cudaMemcpy(tw, htw, sizeof(htw), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&sink, 148 * 512 * 8)); CK(cudaMalloc(&cyc, 8));
    run<0>(tw, sink, cyc, "fwd");
    run<1>(tw, sink, cyc, "fwd-tw2");
    run<2>(tw, sink, cyc, "fp64only");
    return 0;
}
