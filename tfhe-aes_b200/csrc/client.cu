// client.cu — trusted-side harness on the GPU: key generation, encryption and decryption
// (what client.rs:70-175 does through tfhe-rs gen_keys_radix / WopbsKey::new_wopbs_key_only_for_wopbs /
// encrypt_without_padding / decrypt_without_padding).  It exists so that benchmarks, examples and the
// multi-GPU path are self-contained; it is NOT on the server hot path.  The generator is a counter-based
// hash (SplitMix64 finaliser), adequate for tests and benchmarks, not a CSPRNG.
#include <cmath>
#include <cstring>
#include "engine.h"

int prepare_keys_from_device(tfa_ctx *ctx, const u64 *bsk_std_dev);  // engine.cu
int alloc_key_staging(tfa_ctx *ctx);                                   // engine.cu

__host__ __device__ static inline u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ static inline u64 rnd(u64 seed, u64 stream, u64 idx) {
    return mix64(mix64(seed + 0x9E3779B97F4A7C15ull * (stream + 1)) ^ (idx * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull));
}
__device__ static inline u64 gauss_noise(u64 seed, u64 stream, u64 idx, double std_scaled) {
    const double u1 = ((double)(rnd(seed, stream, 2 * idx) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(rnd(seed, stream, 2 * idx + 1) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double g = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    return (u64)__double2ll_rn(g * std_scaled);
}

// ---- LWE encryption: one CTA per ciphertext --------------------------------------------------------
// out[row*out_stride + t] mask, body at column dim.  plaintext per row from `pt` (or computed for the KSK).
__global__ void lwe_encrypt_kernel(const u64 *__restrict__ sk, int dim, u64 seed, u64 stream, double std_scaled, const u64 *__restrict__ pt,
                                   int ksk_levels, int ksk_base_log, const u64 *__restrict__ ksk_in_key, u64 *__restrict__ out, int out_stride) {
    __shared__ u64 red[256];
    const long row = blockIdx.x;
    u64 acc = 0;
    for (int t = threadIdx.x; t < dim; t += blockDim.x) {
        const u64 a = rnd(seed, stream, (u64)row * (dim + 1) + t);
        out[row * out_stride + t] = a;
        acc += a * sk[t];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        u64 m;
        if (ksk_in_key) {  // row = i*levels + (level-1): z_i * q / beta^level
            const long i = row / ksk_levels;
            const int level = (int)(row % ksk_levels) + 1;
            m = ksk_in_key[i] << (64 - ksk_base_log * level);
        } else m = pt[row];
        out[row * out_stride + dim] = red[0] + m + gauss_noise(seed, stream + 1, (u64)row, std_scaled);
    }
}

// ---- GLWE encryption for BSK / PFPKSK: one CTA (512 threads) per ciphertext -------------------------
// mode 0 (BSK):    q = ((i*L + lev)*(k+1) + row); message row<k: -(s_i<<sh) * S_row ; row==k: +(s_i<<sh)
// mode 1 (PFPKSK): q = ((key*(big+1) + j)*L + lev); scal = (-z_j)<<sh (z_big = -1); key<k: scal*S_key; key==k: -scal
template <int K>
__global__ void __launch_bounds__(512) glwe_keygen_kernel(int mode, const u64 *__restrict__ lwe_sk, const u64 *__restrict__ glwe_sk, int levels,
                                                         int base_log, int big, u64 seed, u64 stream, double std_scaled, u64 *__restrict__ out) {
    __shared__ u64 A[K][512];
    __shared__ uint8_t S[K][512];
    const long q = blockIdx.x;
    const int i = threadIdx.x;
    u64 *ct = out + (size_t)q * (K + 1) * 512;
    for (int r = 0; r < K; r++) {
        const u64 a = rnd(seed, stream, ((u64)q * (K + 1) + r) * 512 + i);
        A[r][i] = a;
        ct[(size_t)r * 512 + i] = a;
        S[r][i] = (uint8_t)glwe_sk[(size_t)r * 512 + i];
    }
    __syncthreads();
    u64 acc = 0;
    for (int r = 0; r < K; r++)
        for (int j = 0; j < 512; j++) {
            const int s = S[r][(i - j) & 511];
            const u64 a = A[r][j];
            if (s) acc += (j <= i) ? a : (u64)0 - a;
        }
    u64 coef;
    int poly;  // K = constant polynomial 1
    if (mode == 0) {
        const int row = (int)(q % (K + 1));
        const long il = q / (K + 1);
        const int lev = (int)(il % levels) + 1;
        const long idx = il / levels;
        const u64 factor = lwe_sk[idx] << (64 - base_log * lev);
        coef = (row < K) ? (u64)0 - factor : factor;
        poly = row;
    } else {
        const int lev = (int)(q % levels) + 1;
        const long kj = q / levels;
        const int j = (int)(kj % (big + 1));
        const int key = (int)(kj / (big + 1));
        const u64 z = (j < big) ? glwe_sk[j] : ~0ull;
        const u64 scal = ((u64)0 - z) << (64 - base_log * lev);
        coef = (key < K) ? scal : (u64)0 - scal;
        poly = key;
    }
    const u64 m = (poly < K) ? coef * (u64)S[poly][i] : (i == 0 ? coef : 0);
    ct[(size_t)K * 512 + i] = acc + m + gauss_noise(seed, stream + 1, (u64)q * 512 + i, std_scaled);
}

// phase of big-key LWEs -> decoded bit (decrypt_without_padding, client.rs:154)
__global__ void lwe_decrypt_bits_kernel(const u64 *__restrict__ sk, int dim, const u64 *__restrict__ ct, uint8_t *__restrict__ bits) {
    __shared__ u64 red[256];
    const long row = blockIdx.x;
    const u64 *c = ct + (size_t)row * (dim + 1);
    u64 acc = 0;
    for (int t = threadIdx.x; t < dim; t += blockDim.x) acc += c[t] * sk[t];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) bits[row] = (uint8_t)(((c[dim] - red[0] + (1ull << 62)) >> 63) & 1);
}

extern "C" int tfa_client_keygen(tfa_ctx *ctx, uint64_t seed) {
    RC(tfa_ctx_alloc_keys(ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    RC(alloc_key_staging(ctx));
    const int n = ctx->n, k = ctx->k, big = ctx->big;
    ctx->h_lwe_sk.resize(n); ctx->h_glwe_sk.resize(big);
    for (int i = 0; i < n; i++) ctx->h_lwe_sk[i] = rnd(seed, 1, i) & 1;
    for (int i = 0; i < big; i++) ctx->h_glwe_sk[i] = rnd(seed, 2, i) & 1;
    if (!ctx->d_lwe_sk) CU(cudaMalloc(&ctx->d_lwe_sk, (size_t)n * 8));
    if (!ctx->d_glwe_sk) CU(cudaMalloc(&ctx->d_glwe_sk, (size_t)big * 8));
    CU(cudaMemcpyAsync(ctx->d_lwe_sk, ctx->h_lwe_sk.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_glwe_sk, ctx->h_glwe_sk.data(), (size_t)big * 8, cudaMemcpyHostToDevice, ctx->stream));
    const double two64 = 18446744073709551616.0;
    // KSK
    CU(cudaMemsetAsync(ctx->ksk, 0, ctx->ksk_bytes(), ctx->stream));
    lwe_encrypt_kernel<<<big * ctx->p.ks_level, 256, 0, ctx->stream>>>(ctx->d_lwe_sk, n, seed, 10, ctx->p.lwe_std * two64, nullptr, ctx->p.ks_level,
                                                                     ctx->p.ks_base_log, ctx->d_glwe_sk, ctx->ksk, ctx->ks_cols_pad);
    CU(cudaGetLastError());
    // BSK (standard domain, temporary) and PFPKSK
    u64 *bsk_std = nullptr;
    const long nbsk = (long)n * ctx->p.pbs_level * (k + 1);
    CU(cudaMalloc(&bsk_std, (size_t)nbsk * ctx->gsz * 8));
    const long npf = (long)(k + 1) * (big + 1) * ctx->p.pfks_level;
    if (k == 4) {
        glwe_keygen_kernel<4><<<(unsigned)nbsk, 512, 0, ctx->stream>>>(0, ctx->d_lwe_sk, ctx->d_glwe_sk, ctx->p.pbs_level, ctx->p.pbs_base_log, big, seed, 20,
                                                                      ctx->p.glwe_std * two64, bsk_std);
        glwe_keygen_kernel<4><<<(unsigned)npf, 512, 0, ctx->stream>>>(1, ctx->d_lwe_sk, ctx->d_glwe_sk, ctx->p.pfks_level, ctx->p.pfks_base_log, big, seed, 30,
                                                                     ctx->p.pfks_std * two64, ctx->pfpksk);
    } else {
        glwe_keygen_kernel<1><<<(unsigned)nbsk, 512, 0, ctx->stream>>>(0, ctx->d_lwe_sk, ctx->d_glwe_sk, ctx->p.pbs_level, ctx->p.pbs_base_log, big, seed, 20,
                                                                      ctx->p.glwe_std * two64, bsk_std);
        glwe_keygen_kernel<1><<<(unsigned)npf, 512, 0, ctx->stream>>>(1, ctx->d_lwe_sk, ctx->d_glwe_sk, ctx->p.pfks_level, ctx->p.pfks_base_log, big, seed, 30,
                                                                     ctx->p.pfks_std * two64, ctx->pfpksk);
    }
    CU(cudaGetLastError());
    ctx->launches += 3;
    int rc = prepare_keys_from_device(ctx, bsk_std);
    cudaFree(bsk_std);
    return rc;
}

static int encrypt_bytes_dev_nolock(tfa_ctx *ctx, const uint8_t *bytes, int count, uint64_t seed, u64 *out_dev) {
    if (!ctx->d_glwe_sk) return ctx->fail(TFA_ERR_STATE, "client keys not generated (tfa_client_keygen)");
    std::vector<u64> pt((size_t)count * 8);
    for (int i = 0; i < count; i++) for (int j = 0; j < 8; j++) pt[(size_t)i * 8 + j] = (u64)((bytes[i] >> j) & 1) << 63;  // client.rs:126-138
    u64 *d_pt = nullptr;
    CU(cudaMalloc(&d_pt, pt.size() * 8));
    CU(cudaMemcpyAsync(d_pt, pt.data(), pt.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    lwe_encrypt_kernel<<<count * 8, 256, 0, ctx->stream>>>(ctx->d_glwe_sk, ctx->big, seed, 40, ctx->p.glwe_std * 18446744073709551616.0, d_pt, 0, 0, nullptr,
                                                          out_dev, ctx->lw);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_pt);
    if (e != cudaSuccess) return ctx->fail_cuda("lwe_encrypt_kernel", e);
    ctx->launches++;
    return TFA_OK;
}
extern "C" int tfa_client_encrypt_bytes_dev(tfa_ctx *ctx, const uint8_t *bytes, int count, uint64_t seed, uint64_t *out_dev) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    return encrypt_bytes_dev_nolock(ctx, bytes, count, seed, out_dev);
}
extern "C" int tfa_client_encrypt_bytes(tfa_ctx *ctx, const uint8_t *bytes, int count, uint64_t seed, uint64_t *out) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    u64 *d = nullptr;
    const size_t w = (size_t)count * ctx->byte_words();
    CU(cudaMalloc(&d, w * 8));
    int rc = encrypt_bytes_dev_nolock(ctx, bytes, count, seed, d);
    if (rc == TFA_OK) {
        cudaError_t e = cudaMemcpy(out, d, w * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = ctx->fail_cuda("cudaMemcpy", e);
    }
    cudaFree(d);
    return rc;
}
static int decrypt_bytes_dev_nolock(tfa_ctx *ctx, const u64 *ct_dev, int count, uint8_t *bytes) {
    if (!ctx->d_glwe_sk) return ctx->fail(TFA_ERR_STATE, "client keys not generated (tfa_client_keygen)");
    uint8_t *d_bits = nullptr;
    CU(cudaMalloc(&d_bits, (size_t)count * 8));
    lwe_decrypt_bits_kernel<<<count * 8, 256, 0, ctx->stream>>>(ctx->d_glwe_sk, ctx->big, ct_dev, d_bits);
    std::vector<uint8_t> bits((size_t)count * 8);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(bits.data(), d_bits, bits.size(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_bits);
    if (e != cudaSuccess) return ctx->fail_cuda("lwe_decrypt_bits_kernel", e);
    ctx->launches++;
    for (int i = 0; i < count; i++) { uint8_t v = 0; for (int j = 0; j < 8; j++) v |= bits[(size_t)i * 8 + j] << j; bytes[i] = v; }
    return TFA_OK;
}
extern "C" int tfa_client_decrypt_bytes_dev(tfa_ctx *ctx, const uint64_t *ct_dev, int count, uint8_t *bytes) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    return decrypt_bytes_dev_nolock(ctx, ct_dev, count, bytes);
}
extern "C" int tfa_client_decrypt_bytes(tfa_ctx *ctx, const uint64_t *ct, int count, uint8_t *bytes) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    u64 *d = nullptr;
    const size_t w = (size_t)count * ctx->byte_words();
    CU(cudaMalloc(&d, w * 8));
    cudaError_t e = cudaMemcpyAsync(d, ct, w * 8, cudaMemcpyHostToDevice, ctx->stream);
    int rc = (e == cudaSuccess) ? decrypt_bytes_dev_nolock(ctx, d, count, bytes) : ctx->fail_cuda("cudaMemcpyAsync", e);
    cudaFree(d);
    return rc;
}
extern "C" int tfa_client_secret_keys(tfa_ctx *ctx, uint64_t *lwe_sk, uint64_t *glwe_sk) {
    if (ctx->h_lwe_sk.empty()) return ctx->fail(TFA_ERR_STATE, "client keys not generated (tfa_client_keygen)");
    memcpy(lwe_sk, ctx->h_lwe_sk.data(), ctx->h_lwe_sk.size() * 8);
    memcpy(glwe_sk, ctx->h_glwe_sk.data(), ctx->h_glwe_sk.size() * 8);
    return TFA_OK;
}
extern "C" int tfa_client_set_secret_keys(tfa_ctx *ctx, const uint64_t *lwe_sk, const uint64_t *glwe_sk) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    ctx->h_lwe_sk.assign(lwe_sk, lwe_sk + ctx->n);
    ctx->h_glwe_sk.assign(glwe_sk, glwe_sk + ctx->big);
    if (!ctx->d_lwe_sk) CU(cudaMalloc(&ctx->d_lwe_sk, (size_t)ctx->n * 8));
    if (!ctx->d_glwe_sk) CU(cudaMalloc(&ctx->d_glwe_sk, (size_t)ctx->big * 8));
    CU(cudaMemcpyAsync(ctx->d_lwe_sk, lwe_sk, (size_t)ctx->n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_glwe_sk, glwe_sk, (size_t)ctx->big * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TFA_OK;
}
