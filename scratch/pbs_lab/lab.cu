// PBS kernel laboratory: times several builds ("variants") of the warp-specialised PBS kernel on one wave of
// ciphertexts with a random Fourier key and checks that every variant reproduces variant 0 bit for bit
// (scheduling changes must not change a single floating-point operation).  Not part of the product.
//   ./lab [count=444] [reps=3]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../tfhe-aes_b200/csrc/kernels.h"
#include "../../tfhe-aes_b200/csrc/twiddle_host.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
typedef cudaError_t (*launch_fn)(int, int, int, int, const PbsArgs &, cudaStream_t);
struct Variant { const char *name; launch_fn fn; int G; };
#define DECL(v) cudaError_t launch_##v(int, int, int, int, const PbsArgs &, cudaStream_t);
#include "variants.inc"
#undef DECL
static Variant variants[] = {
#define DECL(v) {#v, launch_##v, 3},
#include "variants.inc"
#undef DECL
};

__global__ void fill_key(double2 *k, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + seed;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        uint64_t y = x * 0xD6E8FEB86659FD93ull + 12345; y ^= y >> 32;
        // Fourier coefficients of a torus polynomial: magnitude ~ 2^63 * sqrt(512) / 2
        k[i] = make_double2((double)(int64_t)x * 8.0, (double)(int64_t)y * 8.0);
    }
}
__global__ void fill_u64(uint64_t *p, size_t n, uint64_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + seed;
        x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
        p[i] = x;
    }
}

int main(int argc, char **argv) {
    const int count = argc > 1 ? atoi(argv[1]) : 444, reps = argc > 2 ? atoi(argv[2]) : 3;
    const int n = argc > 3 ? atoi(argv[3]) : 669, K = 4, N = 512;
    const size_t key_elems = (size_t)n * 5 * 5 * 5 * 256;
    double2 *bsk, *tw; uint64_t *lwe, *lut, *out, *ref;
    CK(cudaMalloc(&bsk, key_elems * sizeof(double2)));
    CK(cudaMalloc(&tw, 256 * sizeof(double2)));
    CK(cudaMalloc(&lwe, (size_t)count * (n + 1) * 8));
    CK(cudaMalloc(&lut, N * 8));
    const size_t out_words = (size_t)count * (K * N + 1);
    CK(cudaMalloc(&out, out_words * 8));
    CK(cudaMalloc(&ref, out_words * 8));
    fill_key<<<1024, 256>>>(bsk, key_elems, 1);
    fill_u64<<<256, 256>>>(lwe, (size_t)count * (n + 1), 2);
    fill_u64<<<1, 256>>>(lut, N, 3);
    cd htw[256]; make_twiddle_table(htw);
    CK(cudaMemcpy(tw, htw, sizeof(htw), cudaMemcpyHostToDevice));
    CK(cudaDeviceSynchronize());
    PbsArgs a{};
    a.lwe_in = lwe; a.bsk = bsk; a.tw = tw; a.lut = lut; a.in_scale = 1; a.pre_add_body = 1ull << 62; a.post_add = 1ull << 48;
    a.lwe_dim = n; a.count = count; a.dbg = nullptr;
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<uint64_t> href(out_words), hout(out_words);
    const int nv = sizeof(variants) / sizeof(variants[0]);
    const char *only = getenv("LAB_ONLY");
    if (getenv("LAB_G")) for (int v = 0; v < nv; v++) variants[v].G = atoi(getenv("LAB_G"));   // ciphertexts per CTA (pbs_ws_kernel variants: 1, 2, 3)
    for (int v = 0; v < nv; v++) {
        if (only && v > 0 && !strstr(only, variants[v].name)) continue;
        a.out = v == 0 ? ref : out;
        CK(cudaMemsetAsync(a.out, 0xEE, out_words * 8, s));
        float best = 1e30f;
        for (int r = 0; r < reps + 1; r++) {
            CK(cudaEventRecord(e0, s));
            cudaError_t e = variants[v].fn(K, variants[v].G, 8, 5, a, s);
            if (e != cudaSuccess) { printf("%-28s launch failed: %s\n", variants[v].name, cudaGetErrorString(e)); break; }
            CK(cudaEventRecord(e1, s));
            e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { printf("%-28s run failed: %s\n", variants[v].name, cudaGetErrorString(e)); return 1; }
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        size_t bad = 0;
        if (v == 0) CK(cudaMemcpy(href.data(), ref, out_words * 8, cudaMemcpyDeviceToHost));
        else {
            CK(cudaMemcpy(hout.data(), out, out_words * 8, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < out_words; i++) bad += hout[i] != href[i];
        }
        if (getenv("LAB_PHASES")) {   // per-activity cycle counters of CTA 0 (TIMING instantiation of the kernel)
            uint64_t *dbg; CK(cudaMalloc(&dbg, 4096 * 8)); CK(cudaMemset(dbg, 0, 4096 * 8));
            PbsArgs b = a; b.dbg = dbg;
            if (variants[v].fn(K, variants[v].G, 8, 5, b, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess) {
                uint64_t h[9]; CK(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
                static const char *nm[9] = {"F.decomp", "F.pass1", "F.waitslot", "F.post", "F.waitinv", "F.inv", "M.wait", "M.mac", "M.handover"};
                printf("   phases/step:");
                for (int k = 0; k < 9; k++) printf(" %s=%.0f", nm[k], (double)h[k] / n);
                printf("\n");
            } else { printf("   timing launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); }
            CK(cudaFree(dbg));
        }
        printf("%-28s %8.3f ms  %8.0f PBS/s  %6.2f TFLOP/s  mismatching words %zu%s\n", variants[v].name, best, count / best * 1e3,
               count * 407608320.0 * (n / 669.0) / best * 1e-9, bad, v == 0 ? " (reference)" : "");
        fflush(stdout);
    }
    return 0;
}
