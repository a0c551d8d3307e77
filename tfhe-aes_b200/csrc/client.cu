// client.cu — trusted-side harness on the GPU: key generation, encryption and decryption
// (what client.rs:70-175 does through tfhe-rs gen_keys_radix / WopbsKey::new_wopbs_key_only_for_wopbs /
// encrypt_without_padding / decrypt_without_padding).  It exists so that benchmarks, examples and the
// multi-GPU path are self-contained; it is NOT on the server hot path.
//
// Randomness: every mask word, secret-key bit and noise sample is a word of the ChaCha20 key stream (the 64-bit
// counter / 64-bit nonce variant) under a 256-bit key: word idx of stream s = word idx mod 8 of block idx / 8 with
// nonce s, so any word is computed independently by the thread that needs it (counter-based, no state).  With
// seed = 0 the key comes from the operating system (getrandom) and is fresh for every key generation and for every
// encryption call, as tfhe-rs seeds its CSPRNG (client.rs:106, :128).  A non-zero seed derives the key from the seed
// alone: reproducible material for tests and benchmarks, INSECURE (anyone who knows the seed regenerates the keys).
#include <cmath>
#include <cstring>
#include <sys/random.h>
#include "engine.h"

int prepare_keys_from_device(tfa_ctx *ctx, const u64 *bsk_std_dev);  // engine.cu
int alloc_key_staging(tfa_ctx *ctx);                                   // engine.cu

struct RngKey { uint32_t k[8]; };
__host__ __device__ static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define CHACHA_QR(a, b, c, d) a += b; d = rotl32(d ^ a, 16); c += d; b = rotl32(b ^ c, 12); a += b; d = rotl32(d ^ a, 8); c += d; b = rotl32(b ^ c, 7)
// words 2w, 2w+1 of ChaCha20 block `counter` under (key, nonce) as one u64
__host__ __device__ static inline u64 chacha20_word(const RngKey &key, u64 nonce, u64 counter, int w) {
    uint32_t x[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5], key.k[6], key.k[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)nonce, (uint32_t)(nonce >> 32)};
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = x[i];
#pragma unroll
    for (int r = 0; r < 10; r++) {
        CHACHA_QR(x[0], x[4], x[8], x[12]); CHACHA_QR(x[1], x[5], x[9], x[13]); CHACHA_QR(x[2], x[6], x[10], x[14]); CHACHA_QR(x[3], x[7], x[11], x[15]);
        CHACHA_QR(x[0], x[5], x[10], x[15]); CHACHA_QR(x[1], x[6], x[11], x[12]); CHACHA_QR(x[2], x[7], x[8], x[13]); CHACHA_QR(x[3], x[4], x[9], x[14]);
    }
    u64 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)      // static indexing only (keeps x[] in registers); w is in 0..7
        if (i == w) { lo = x[2 * i] + s[2 * i]; hi = x[2 * i + 1] + s[2 * i + 1]; }
    return lo | (hi << 32);
}
__host__ __device__ static inline u64 rnd(const RngKey &key, u64 stream, u64 idx) { return chacha20_word(key, stream, idx >> 3, (int)(idx & 7)); }
__device__ static inline u64 gauss_noise(const RngKey &key, u64 stream, u64 idx, double std_scaled) {
    const double u1 = ((double)(rnd(key, stream, 2 * idx) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(rnd(key, stream, 2 * idx + 1) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double g = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    return (u64)__double2ll_rn(g * std_scaled);
}
// seed = 0: 256 bits from the operating system; otherwise a key that is a function of the seed only (tests, benchmarks)
static int make_rng_key(tfa_ctx *ctx, uint64_t seed, RngKey *key) {
    if (seed == 0) {
        size_t got = 0;
        while (got < sizeof(key->k)) {
            ssize_t r = getrandom((char *)key->k + got, sizeof(key->k) - got, 0);
            if (r < 0) return ctx->fail(TFA_ERR_STATE, "getrandom failed: no entropy for key generation / encryption");
            got += (size_t)r;
        }
        return TFA_OK;
    }
    const RngKey base = {{0x74666131u, 0x2d746573u, 0x742d6b65u, 0x79000000u, (uint32_t)seed, (uint32_t)(seed >> 32), 0x5eed5eedu, 0x0badc0deu}};
    for (int i = 0; i < 4; i++) {
        const u64 w = chacha20_word(base, 0x7365656473ull, 0, i);
        key->k[2 * i] = (uint32_t)w; key->k[2 * i + 1] = (uint32_t)(w >> 32);
    }
    return TFA_OK;
}

// the generator's block function on the host, so that the CPU test-suite can pin it to the RFC 7539 §2.3.2 vector
extern "C" void tfa_rng_block(const uint32_t key[8], uint64_t nonce, uint64_t counter, uint64_t out[8]) {
    RngKey k;
    memcpy(k.k, key, sizeof(k.k));
    for (int w = 0; w < 8; w++) out[w] = chacha20_word(k, nonce, counter, w);
}

// ---- LWE encryption: one CTA per ciphertext --------------------------------------------------------
// out[row*out_stride + t] mask, body at column dim.  plaintext per row from `pt` (or computed for the KSK).
__global__ void lwe_encrypt_kernel(const u64 *__restrict__ sk, int dim, RngKey seed, u64 stream, double std_scaled, const u64 *__restrict__ pt,
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
                                                         int base_log, int big, RngKey seed, u64 stream, double std_scaled, u64 *__restrict__ out) {
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

extern "C" int tfa_client_keygen(tfa_ctx *ctx, uint64_t seed_u64) {
    RC(tfa_ctx_alloc_keys(ctx));
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    RC(alloc_key_staging(ctx));
    const int n = ctx->n, k = ctx->k, big = ctx->big;
    RngKey seed;
    RC(make_rng_key(ctx, seed_u64, &seed));
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

static int encrypt_bytes_dev_nolock(tfa_ctx *ctx, const uint8_t *bytes, int count, uint64_t seed_u64, u64 *out_dev) {
    if (!ctx->d_glwe_sk) return ctx->fail(TFA_ERR_STATE, "client keys not generated (tfa_client_keygen)");
    RngKey seed;
    RC(make_rng_key(ctx, seed_u64, &seed));
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
