// tfhe_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see tfhe_oracle.h header comment).
//
// Restates, function by function, what the reference executes for one encrypted byte
// (SURVEY.md §9).  The reference's own files give the call sites and arguments; the arithmetic
// lives in the un-vendored crate tfhe 0.11.2 / tfhe-fft 0.7.0 and is restated from its published
// algorithms.  Ciphertext-level parity with real tfhe-rs is UNPINNED (no such vector exists in the
// reference); plaintext-level parity is pinned by FIPS-197 / SP 800-38A known answers.
//
// Build: see oracle/Makefile  (g++ -O3 -fopenmp -shared -fPIC).
#include "tfhe_oracle.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef int64_t i64;
typedef std::complex<double> cplx;

// ------------------------------------------------------------------------------------------------
// deterministic RNG (oracle-local; tfhe-rs seeds tfhe-csprng from OS entropy, so nothing to match)
// ------------------------------------------------------------------------------------------------
namespace {
struct Rng {
    u64 s[4];
    static u64 splitmix(u64 &x) {
        u64 z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(u64 seed, u64 stream = 0) {
        u64 x = seed ^ (stream * 0xD6E8FEB86659FD93ull + 0x1234567ull);
        for (int i = 0; i < 4; i++) s[i] = splitmix(x);
    }
    static u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
    u64 next() {
        u64 r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return ((next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    double gauss() {  // Box-Muller
        double u1 = uniform(), u2 = uniform();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
    }
    u64 noise(double std) {  // torus noise as u64
        double v = gauss() * std * 18446744073709551616.0;
        return (u64)(i64)std::llround(v);
    }
};
}  // namespace

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct orc_ctx {
    orc_params p;
    int n, k, N, M, big;  // big = k*N
    std::vector<cplx> twist;   // [M]  exp(i*pi*j/N)
    std::vector<cplx> wtab;    // [M/2] exp(2*pi*i*j/M)
    std::vector<double> wstage_fwd, wstage_inv;   // per-stage contiguous copies of wtab (and conjugates), interleaved re/im
    std::vector<int> bitrev;   // [M]
    std::vector<u64> lwe_sk, glwe_sk, bsk, ksk, pfpksk;
    std::vector<double> bsk_f;  // [n][l][k+1][k+1][M][2]
    Rng enc_rng{0x1234, 77};
};

// ------------------------------------------------------------------------------------------------
// SURVEY §9.3 — signed decomposition (tfhe-rs SignedDecomposer: closest_representable +
// decompose_one_level).  Digits come out from level `level` (least significant) to level 1.
// ------------------------------------------------------------------------------------------------
static inline u64 decomp_init_state(u64 x, int base_log, int level) {
    int r = 64 - base_log * level;
    u64 v = (x >> r) + ((x >> (r - 1)) & 1);  // closest representable, expressed in units of 2^r
    if (base_log * level < 64) v &= (~0ull) >> r;  // the << r of the reference wraps the top carry away
    return v;
}
static inline i64 decomp_next(u64 &state, int base_log) {
    u64 mask = (1ull << base_log) - 1;
    u64 res = state & mask;
    state >>= base_log;
    u64 carry = ((res - 1) | state) & res;
    carry >>= (base_log - 1);
    state += carry;
    return (i64)(res - (carry << base_log));
}
extern "C" void orc_decompose(uint64_t x, int base_log, int level, int64_t *digits) {
    u64 st = decomp_init_state(x, base_log, level);
    for (int i = 0; i < level; i++) digits[i] = decomp_next(st, base_log);
}

// ------------------------------------------------------------------------------------------------
// SURVEY §9.6 — negacyclic FFT, tfhe-fft convention: z_j = p_j + i p_{j+N/2}, twist exp(i*pi*j/N),
// N/2-point complex DFT.  (tfhe-rs additionally scales torus inputs by 2^-64 and back by 2^64;
// both are exact power-of-two scalings, omitted here.)
// ------------------------------------------------------------------------------------------------
static void fft_core(const orc_ctx *c, cplx *a, bool inverse) {
    // iterative radix-2, decimation in time.  Same operations per element as the textbook loop over std::complex
    // (product = (ac - bd, ad + bc), then sum / difference); written on the interleaved doubles with per-stage contiguous
    // twiddles so that the compiler vectorises the butterflies — this is also the CPU arm of bench.py.
    const int M = c->M;
    for (int i = 0; i < M; i++) {
        int j = c->bitrev[i];
        if (i < j) std::swap(a[i], a[j]);
    }
    double *p = reinterpret_cast<double *>(a);
    const double *wt = inverse ? c->wstage_inv.data() : c->wstage_fwd.data();
    for (int len = 2; len <= M; len <<= 1) {
        const int half = len >> 1;
        const double *w = wt + 2 * (half - 1);          // stage tables are concatenated: offsets 0, 1, 3, 7, ...
        for (int i = 0; i < M; i += len) {
            double *u = p + 2 * i, *v = p + 2 * (i + half);
#pragma omp simd
            for (int j = 0; j < half; j++) {
                const double br = v[2 * j], bi = v[2 * j + 1], wr = w[2 * j], wi = w[2 * j + 1];
                const double tr = br * wr - bi * wi, ti = br * wi + bi * wr;
                const double ur = u[2 * j], ui = u[2 * j + 1];
                u[2 * j] = ur + tr; u[2 * j + 1] = ui + ti;
                v[2 * j] = ur - tr; v[2 * j + 1] = ui - ti;
            }
        }
    }
}
static void fft_forward_signed(const orc_ctx *c, const i64 *re, const i64 *im, cplx *out) {
    for (int j = 0; j < c->M; j++) out[j] = cplx((double)re[j], (double)im[j]) * c->twist[j];
    fft_core(c, out, false);
}
static inline u64 f64_to_torus(double v) {  // round to nearest integer, reduce mod 2^64
    double r = v - 18446744073709551616.0 * std::nearbyint(v * (1.0 / 18446744073709551616.0));
    double rr = std::nearbyint(r);
    if (rr >= 9223372036854775808.0) rr -= 18446744073709551616.0;
    return (u64)(i64)rr;
}
static void fft_add_backward(const orc_ctx *c, cplx *f, u64 *poly) {  // destroys f
    const int M = c->M;
    fft_core(c, f, true);
    const double inv = 1.0 / M;
    for (int j = 0; j < M; j++) {
        cplx z = f[j] * std::conj(c->twist[j]) * inv;
        poly[j] += f64_to_torus(z.real());
        poly[j + M] += f64_to_torus(z.imag());
    }
}
extern "C" void orc_fft_forward_torus(const orc_ctx *c, const uint64_t *poly, double *out) {
    fft_forward_signed(c, (const i64 *)poly, (const i64 *)poly + c->M, (cplx *)out);
}
extern "C" void orc_fft_forward_integer(const orc_ctx *c, const int64_t *poly, double *out) {
    fft_forward_signed(c, poly, poly + c->M, (cplx *)out);
}
extern "C" void orc_fft_add_backward_torus(const orc_ctx *c, const double *fourier, uint64_t *poly) {
    std::vector<cplx> tmp((const cplx *)fourier, (const cplx *)fourier + c->M);
    fft_add_backward(c, tmp.data(), poly);
}

// ------------------------------------------------------------------------------------------------
// context creation
// ------------------------------------------------------------------------------------------------
extern "C" orc_ctx *orc_create(const orc_params *p) {
    orc_ctx *c = new orc_ctx();
    c->p = *p;
    c->n = p->lwe_dim; c->k = p->glwe_dim; c->N = p->poly_size; c->M = c->N / 2; c->big = c->k * c->N;
    c->twist.resize(c->M); c->wtab.resize(c->M / 2); c->bitrev.resize(c->M);
    for (int j = 0; j < c->M; j++) c->twist[j] = std::polar(1.0, M_PI * j / c->N);
    for (int j = 0; j < c->M / 2; j++) c->wtab[j] = std::polar(1.0, 2.0 * M_PI * j / c->M);
    for (int len = 2; len <= c->M; len <<= 1)
        for (int j = 0; j < len / 2; j++) {
            const cplx w = c->wtab[(size_t)j * (c->M / len)];
            c->wstage_fwd.push_back(w.real()); c->wstage_fwd.push_back(w.imag());
            c->wstage_inv.push_back(w.real()); c->wstage_inv.push_back(-w.imag());
        }
    int lg = 0; while ((1 << lg) < c->M) lg++;
    for (int i = 0; i < c->M; i++) { int r = 0; for (int b = 0; b < lg; b++) if (i >> b & 1) r |= 1 << (lg - 1 - b); c->bitrev[i] = r; }
    return c;
}
extern "C" void orc_destroy(orc_ctx *c) { delete c; }
extern "C" void orc_set_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}
extern "C" int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
extern "C" const uint64_t *orc_lwe_sk(const orc_ctx *c) { return c->lwe_sk.data(); }
extern "C" const uint64_t *orc_glwe_sk(const orc_ctx *c) { return c->glwe_sk.data(); }
extern "C" const uint64_t *orc_bsk(const orc_ctx *c) { return c->bsk.data(); }
extern "C" const uint64_t *orc_ksk(const orc_ctx *c) { return c->ksk.data(); }
extern "C" const uint64_t *orc_pfpksk(const orc_ctx *c) { return c->pfpksk.data(); }

// ------------------------------------------------------------------------------------------------
// encryption helpers
// ------------------------------------------------------------------------------------------------
// body += a * s (negacyclic), s binary
static void negacyclic_mul_binary_add(u64 *body, const u64 *a, const u64 *s, int N) {
    for (int j = 0; j < N; j++) {
        if (!s[j]) continue;
        for (int i = 0; i < N - j; i++) body[i + j] += a[i];
        for (int i = N - j; i < N; i++) body[i + j - N] -= a[i];
    }
}
// GLWE encryption of plaintext polynomial pt (added into body) under glwe_sk
static void glwe_encrypt(const orc_ctx *c, Rng &rng, const u64 *pt, double std, u64 *out) {
    const int N = c->N, k = c->k;
    u64 *body = out + (size_t)k * N;
    for (int i = 0; i < N; i++) body[i] = pt[i] + rng.noise(std);
    for (int r = 0; r < k; r++) {
        u64 *a = out + (size_t)r * N;
        for (int i = 0; i < N; i++) a[i] = rng.next();
        negacyclic_mul_binary_add(body, a, c->glwe_sk.data() + (size_t)r * N, N);
    }
}
static void lwe_encrypt(const u64 *sk, int dim, Rng &rng, u64 pt, double std, u64 *out) {
    u64 b = pt + rng.noise(std);
    for (int i = 0; i < dim; i++) { out[i] = rng.next(); b += out[i] * sk[i]; }
    out[dim] = b;
}
static u64 lwe_phase(const u64 *sk, int dim, const u64 *ct) {
    u64 b = ct[dim];
    for (int i = 0; i < dim; i++) b -= ct[i] * sk[i];
    return b;
}

// ------------------------------------------------------------------------------------------------
// keygen — client.rs:106-107 (gen_keys_radix + WopbsKey::new_wopbs_key_only_for_wopbs).
// One BSK and one KSK serve the whole chain; PFPKSK list per
// tfhe-rs generate_circuit_bootstrap_lwe_pfpksk_list: f(x) = -x, polynomials S_0..S_{k-1}, then -1.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_keygen(orc_ctx *c, uint64_t seed) {
    const int n = c->n, k = c->k, N = c->N, big = c->big;
    const orc_params &p = c->p;
    Rng kr(seed, 1);
    c->lwe_sk.resize(n); c->glwe_sk.resize(big);
    for (auto &b : c->lwe_sk) b = kr.next() & 1;
    for (auto &b : c->glwe_sk) b = kr.next() & 1;

    // KSK: big -> small, [i][level-1][n+1] encrypts z_i * q/beta^level under the small key, noise lwe_std
    c->ksk.assign((size_t)big * p.ks_level * (n + 1), 0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < big; i++) {
        Rng r(seed, 0x10000000ull + i);
        for (int l = 1; l <= (int)p.ks_level; l++) {
            u64 pt = c->glwe_sk[i] << (64 - p.ks_base_log * l);
            lwe_encrypt(c->lwe_sk.data(), n, r, pt, p.lwe_std, &c->ksk[((size_t)i * p.ks_level + (l - 1)) * (n + 1)]);
        }
    }
    // BSK: [i][level-1][row][ (k+1)N ]; row r<k: -s_i q/beta^l * S_r ; row k: +s_i q/beta^l
    const size_t glwe_sz = (size_t)(k + 1) * N;
    c->bsk.assign((size_t)n * p.pbs_level * (k + 1) * glwe_sz, 0);
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < n; i++) {
        Rng r(seed, 0x20000000ull + i);
        std::vector<u64> pt(N);
        for (int l = 1; l <= (int)p.pbs_level; l++) {
            u64 factor = c->lwe_sk[i] << (64 - p.pbs_base_log * l);
            for (int row = 0; row <= k; row++) {
                if (row < k) for (int j = 0; j < N; j++) pt[j] = (u64)0 - factor * c->glwe_sk[(size_t)row * N + j];
                else { std::fill(pt.begin(), pt.end(), 0); pt[0] = factor; }
                glwe_encrypt(c, r, pt.data(), p.glwe_std, &c->bsk[(((size_t)i * p.pbs_level + (l - 1)) * (k + 1) + row) * glwe_sz]);
            }
        }
    }
    // PFPKSK list: [key r][j in 0..big][level-1][(k+1)N], message f(1)*z_j*q/beta^l * P_r,
    // f(1) = -1, z_big = -1, P_r = S_r (r<k) or the constant polynomial -1 (r=k).
    c->pfpksk.assign((size_t)(k + 1) * (big + 1) * p.pfks_level * glwe_sz, 0);
#pragma omp parallel for schedule(dynamic, 16) collapse(2)
    for (int key = 0; key <= k; key++)
        for (int j = 0; j <= big; j++) {
            Rng r(seed, 0x30000000ull + (u64)key * 0x100000 + j);
            std::vector<u64> pt(N);
            u64 z = (j < big) ? c->glwe_sk[j] : ~0ull;  // -1 for the body element
            u64 fz = (u64)0 - z;                         // f(1) * z with f(x) = -x
            for (int l = 1; l <= (int)p.pfks_level; l++) {
                u64 scal = fz << (64 - p.pfks_base_log * l);
                if (key < k) for (int t = 0; t < N; t++) pt[t] = scal * c->glwe_sk[(size_t)key * N + t];
                else { std::fill(pt.begin(), pt.end(), 0); pt[0] = (u64)0 - scal; }
                glwe_encrypt(c, r, pt.data(), p.pfks_std,
                             &c->pfpksk[((((size_t)key * (big + 1)) + j) * p.pfks_level + (l - 1)) * glwe_sz]);
            }
        }
    // Fourier BSK
    const int M = c->M;
    const size_t npoly = (size_t)n * p.pbs_level * (k + 1) * (k + 1);
    c->bsk_f.assign(npoly * M * 2, 0.0);
#pragma omp parallel for schedule(static)
    for (long q = 0; q < (long)npoly; q++)
        orc_fft_forward_torus(c, &c->bsk[(size_t)q * N], &c->bsk_f[(size_t)q * M * 2]);
    c->enc_rng = Rng(seed, 99);
}

// ------------------------------------------------------------------------------------------------
// client encodings — client.rs:126-138 (encrypt_without_padding, MSB byte first is the caller's
// job), client.rs:147-175 (decrypt_without_padding).  Block j of a byte = bit j (LSB first).
// EncryptionKeyChoice::Big => big key, glwe noise.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_seed_encryption(orc_ctx *c, uint64_t seed) { c->enc_rng = Rng(seed, 99); }
extern "C" void orc_encrypt_bits(orc_ctx *c, const uint8_t *bits, int count, uint64_t *out) {
    for (int i = 0; i < count; i++)
        lwe_encrypt(c->glwe_sk.data(), c->big, c->enc_rng, (u64)(bits[i] & 1) << 63, c->p.glwe_std, out + (size_t)i * (c->big + 1));
}
extern "C" void orc_encrypt_bytes(orc_ctx *c, const uint8_t *bytes, int count, uint64_t *out) {
    std::vector<uint8_t> bits((size_t)count * 8);
    for (int i = 0; i < count; i++) for (int j = 0; j < 8; j++) bits[(size_t)i * 8 + j] = (bytes[i] >> j) & 1;
    orc_encrypt_bits(c, bits.data(), count * 8, out);
}
extern "C" void orc_trivial_bytes(const orc_ctx *c, const uint8_t *bytes, int count, uint64_t *out) {
    size_t sz = c->big + 1;
    memset(out, 0, (size_t)count * 8 * sz * 8);
    for (int i = 0; i < count; i++) for (int j = 0; j < 8; j++) out[((size_t)i * 8 + j) * sz + c->big] = (u64)((bytes[i] >> j) & 1) << 63;
}
extern "C" void orc_phase_big(const orc_ctx *c, const uint64_t *ct, int count, uint64_t *phase) {
    for (int i = 0; i < count; i++) phase[i] = lwe_phase(c->glwe_sk.data(), c->big, ct + (size_t)i * (c->big + 1));
}
extern "C" void orc_phase_small(const orc_ctx *c, const uint64_t *ct, int count, uint64_t *phase) {
    for (int i = 0; i < count; i++) phase[i] = lwe_phase(c->lwe_sk.data(), c->n, ct + (size_t)i * (c->n + 1));
}
extern "C" void orc_decrypt_bits(const orc_ctx *c, const uint64_t *ct, int count, uint8_t *bits, int64_t *err) {
    for (int i = 0; i < count; i++) {
        u64 ph = lwe_phase(c->glwe_sk.data(), c->big, ct + (size_t)i * (c->big + 1));
        u64 b = ((ph + (1ull << 62)) >> 63) & 1;
        bits[i] = (uint8_t)b;
        if (err) err[i] = (i64)(ph - (b << 63));
    }
}
extern "C" void orc_decrypt_bytes(const orc_ctx *c, const uint64_t *ct, int count, uint8_t *bytes) {
    std::vector<uint8_t> bits((size_t)count * 8);
    orc_decrypt_bits(c, ct, count * 8, bits.data(), nullptr);
    for (int i = 0; i < count; i++) { uint8_t v = 0; for (int j = 0; j < 8; j++) v |= bits[(size_t)i * 8 + j] << j; bytes[i] = v; }
}
extern "C" void orc_encrypt_lwe_small(orc_ctx *c, const uint64_t *plain, int count, uint64_t *out) {
    for (int i = 0; i < count; i++) lwe_encrypt(c->lwe_sk.data(), c->n, c->enc_rng, plain[i], c->p.lwe_std, out + (size_t)i * (c->n + 1));
}
extern "C" void orc_glwe_phase(const orc_ctx *c, const uint64_t *glwe, int count, uint64_t *phase) {
    const int N = c->N, k = c->k;
    for (int g = 0; g < count; g++) {
        const u64 *ct = glwe + (size_t)g * (k + 1) * N;
        u64 *ph = phase + (size_t)g * N;
        std::vector<u64> acc(N, 0);
        for (int r = 0; r < k; r++) negacyclic_mul_binary_add(acc.data(), ct + (size_t)r * N, c->glwe_sk.data() + (size_t)r * N, N);
        for (int i = 0; i < N; i++) ph[i] = ct[(size_t)k * N + i] - acc[i];
    }
}

// ------------------------------------------------------------------------------------------------
// SURVEY §9.4(1) keyswitch — tfhe-rs keyswitch_lwe_ciphertext; call site many_wopbs.rs:194-199
// (inside extract_bits_assign).  out = (0,..,0,b) - sum_i sum_l d_{i,l} KSK[i][l]
// ------------------------------------------------------------------------------------------------
static void keyswitch_one(const orc_ctx *c, const u64 *in, u64 *out) {
    const int n = c->n, big = c->big, L = c->p.ks_level, bl = c->p.ks_base_log;
    std::fill(out, out + n + 1, 0);
    out[n] = in[big];
    for (int i = 0; i < big; i++) {
        u64 st = decomp_init_state(in[i], bl, L);
        for (int l = L; l >= 1; l--) {  // digits come out level L first
            i64 d = decomp_next(st, bl);
            if (!d) continue;
            const u64 *row = &c->ksk[((size_t)i * L + (l - 1)) * (n + 1)];
            for (int t = 0; t <= n; t++) out[t] -= (u64)d * row[t];
        }
    }
}
extern "C" void orc_keyswitch(const orc_ctx *c, const uint64_t *in, int count, uint64_t *out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < count; i++) keyswitch_one(c, in + (size_t)i * (c->big + 1), out + (size_t)i * (c->n + 1));
}

// ------------------------------------------------------------------------------------------------
// SURVEY §9.4(4) external product — tfhe-rs add_external_product_assign.
// ggsw_f: Fourier GGSW [level][row][col][M] complex.  out += ggsw (x) glwe
// ------------------------------------------------------------------------------------------------
static void external_product_add(const orc_ctx *c, const double *ggsw_f, int base_log, int level,
                                 const u64 *glwe, u64 *out) {
    const int N = c->N, M = c->M, k = c->k;
    std::vector<cplx> acc((size_t)(k + 1) * M, cplx(0, 0));
    std::vector<cplx> f(M);
    std::vector<u64> state((size_t)(k + 1) * N);
    std::vector<i64> dig(N);
    for (size_t i = 0; i < state.size(); i++) state[i] = decomp_init_state(glwe[i], base_log, level);
    for (int l = level; l >= 1; l--) {  // level `level` first, matching the decomposition order
        for (int row = 0; row <= k; row++) {
            u64 *st = &state[(size_t)row * N];
            for (int j = 0; j < N; j++) dig[j] = decomp_next(st[j], base_log);
            fft_forward_signed(c, dig.data(), dig.data() + M, f.data());
            const cplx *g = (const cplx *)ggsw_f + ((size_t)(l - 1) * (k + 1) + row) * (k + 1) * M;
            for (int col = 0; col <= k; col++) {
                cplx *a = &acc[(size_t)col * M];
                const cplx *gg = g + (size_t)col * M;
                const double *fp = reinterpret_cast<const double *>(f.data()), *gp = reinterpret_cast<const double *>(gg);
                double *ap = reinterpret_cast<double *>(a);
#pragma omp simd
                for (int j = 0; j < M; j++) {   // a[j] += f[j] * gg[j], same operation order as std::complex
                    const double fr = fp[2 * j], fi = fp[2 * j + 1], gr = gp[2 * j], gi = gp[2 * j + 1];
                    ap[2 * j] += fr * gr - fi * gi;
                    ap[2 * j + 1] += fr * gi + fi * gr;
                }
            }
        }
    }
    for (int col = 0; col <= k; col++) fft_add_backward(c, &acc[(size_t)col * M], out + (size_t)col * N);
}
static void ggsw_to_fourier(const orc_ctx *c, const u64 *ggsw_std, int level, double *out) {
    const size_t npoly = (size_t)level * (c->k + 1) * (c->k + 1);
    for (size_t q = 0; q < npoly; q++) orc_fft_forward_torus(c, ggsw_std + q * c->N, out + q * c->M * 2);
}
extern "C" void orc_external_product_add(const orc_ctx *c, const uint64_t *ggsw_std, int base_log, int level,
                                         const uint64_t *glwe_in, uint64_t *glwe_inout) {
    std::vector<double> f((size_t)level * (c->k + 1) * (c->k + 1) * c->M * 2);
    ggsw_to_fourier(c, ggsw_std, level, f.data());
    external_product_add(c, f.data(), base_log, level, glwe_in, glwe_inout);
}

// polynomial helpers (tfhe-rs polynomial_wrapping_monic_monomial_{mul,div})
static void poly_mul_monomial(const u64 *in, int deg /* in [0, 2N) */, int N, u64 *out) {  // out = in * X^deg
    for (int j = 0; j < N; j++) {
        int t = j + deg;
        bool neg = false;
        while (t >= N) { t -= N; neg = !neg; }
        out[t] = neg ? (u64)0 - in[j] : in[j];
    }
}
static void poly_div_monomial(const u64 *in, int deg, int N, u64 *out) {  // out = in * X^-deg
    int d = ((2 * N - deg) % (2 * N) + 2 * N) % (2 * N);
    poly_mul_monomial(in, d, N, out);
}

// ------------------------------------------------------------------------------------------------
// SURVEY §9.4(3) PBS — tfhe-rs FourierLweBootstrapKey::bootstrap = blind_rotate_assign +
// extract_lwe_sample_from_glwe_ciphertext(.., MonomialDegree(0)); call sites: inside
// circuit_bootstrap_boolean (many_wopbs.rs:253) and the extract_bits loop (many_wopbs.rs:194).
// ------------------------------------------------------------------------------------------------
static inline int mod_switch(u64 a, int log2_2N) { return (int)((a + (1ull << (63 - log2_2N))) >> (64 - log2_2N)); }
static void bootstrap_one(const orc_ctx *c, const u64 *in, const u64 *lut, u64 *out) {
    const int n = c->n, k = c->k, N = c->N;
    int lg = 0; while ((1 << lg) < 2 * N) lg++;
    const size_t gsz = (size_t)(k + 1) * N;
    std::vector<u64> acc(gsz, 0), ct1(gsz);
    int bhat = mod_switch(in[n], lg);
    poly_div_monomial(lut, bhat, N, &acc[(size_t)k * N]);  // mask polys of a trivial GLWE stay 0
    const size_t ggsw_sz = (size_t)c->p.pbs_level * (k + 1) * (k + 1) * c->M * 2;
    for (int i = 0; i < n; i++) {
        if (in[i] == 0) continue;
        int ahat = mod_switch(in[i], lg);
        if (ahat == 0) continue;  // ct1 would be identically 0: the product adds exactly 0
        for (int r = 0; r <= k; r++) {
            poly_mul_monomial(&acc[(size_t)r * N], ahat, N, &ct1[(size_t)r * N]);
            for (int j = 0; j < N; j++) ct1[(size_t)r * N + j] -= acc[(size_t)r * N + j];
        }
        external_product_add(c, &c->bsk_f[(size_t)i * ggsw_sz], c->p.pbs_base_log, c->p.pbs_level, ct1.data(), acc.data());
    }
    // sample extract at coefficient 0
    for (int r = 0; r < k; r++) {
        const u64 *a = &acc[(size_t)r * N];
        out[(size_t)r * N] = a[0];
        for (int j = 1; j < N; j++) out[(size_t)r * N + j] = (u64)0 - a[N - j];
    }
    out[(size_t)k * N] = acc[(size_t)k * N];
}
extern "C" void orc_bootstrap(const orc_ctx *c, const uint64_t *in, int count, const uint64_t *lut, uint64_t *out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < count; i++) bootstrap_one(c, in + (size_t)i * (c->n + 1), lut, out + (size_t)i * (c->big + 1));
}

// ------------------------------------------------------------------------------------------------
// SURVEY §9.4(1) extract_bits — tfhe-rs wop_pbs::extract_bits; call site many_wopbs.rs:194-199.
// Output index 0 = most significant extracted bit.  With ExtractedBitsCount(1) this is one
// keyswitch and zero PBS (the reference's case, many_wopbs.rs:179-199).
// ------------------------------------------------------------------------------------------------
extern "C" void orc_extract_bits(const orc_ctx *c, const uint64_t *in, int delta_log, int nbits, uint64_t *out) {
    const int n = c->n, big = c->big, N = c->N;
    std::vector<u64> buf(in, in + big + 1), shifted(big + 1), ks(n + 1), pbs(big + 1), lut(N);
    for (int bit_idx = 0; bit_idx < nbits; bit_idx++) {
        u64 *dst = out + (size_t)(nbits - 1 - bit_idx) * (n + 1);  // list filled in reverse
        u64 mul = 1ull << (64 - delta_log - bit_idx - 1);
        for (int i = 0; i <= big; i++) shifted[i] = buf[i] * mul;
        keyswitch_one(c, shifted.data(), ks.data());
        memcpy(dst, ks.data(), (size_t)(n + 1) * 8);
        if (bit_idx == nbits - 1) break;
        ks[n] += 1ull << 62;
        for (int j = 0; j < N; j++) lut[j] = (u64)0 - (1ull << (delta_log - 1 + bit_idx));
        bootstrap_one(c, ks.data(), lut.data(), pbs.data());
        pbs[big] += 1ull << (delta_log + bit_idx - 1);
        for (int i = 0; i <= big; i++) buf[i] -= pbs[i];
    }
}

// ------------------------------------------------------------------------------------------------
// SURVEY §9.4(5) PFKS — tfhe-rs private_functional_keyswitch_lwe_ciphertext_into_glwe_ciphertext.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_pfks(const orc_ctx *c, int key, const uint64_t *lwe, uint64_t *glwe) {
    const int big = c->big, L = c->p.pfks_level, bl = c->p.pfks_base_log;
    const size_t gsz = (size_t)(c->k + 1) * c->N;
    std::fill(glwe, glwe + gsz, 0);
    for (int j = 0; j <= big; j++) {  // mask AND body
        u64 st = decomp_init_state(lwe[j], bl, L);
        for (int l = L; l >= 1; l--) {
            i64 d = decomp_next(st, bl);
            if (!d) continue;
            const u64 *row = &c->pfpksk[((((size_t)key * (big + 1)) + j) * L + (l - 1)) * gsz];
            for (size_t t = 0; t < gsz; t++) glwe[t] -= (u64)d * row[t];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SURVEY §9.4(2) circuit_bootstrap_boolean — tfhe-rs wop_pbs::circuit_bootstrap_boolean +
// homomorphic_shift_boolean; call site many_wopbs.rs:253-261 with DeltaLog(63).
// ggsw_out: [cbs_level][k+1 rows][(k+1)N]; level matrix index 0 = level 1.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_circuit_bootstrap_boolean(const orc_ctx *c, const uint64_t *lwe_in, int delta_log, uint64_t *ggsw_out) {
    const int n = c->n, k = c->k, N = c->N, big = c->big;
    const size_t gsz = (size_t)(k + 1) * N;
    std::vector<u64> shifted(n + 1), lut(N), bs(big + 1);
    for (int lvl = 1; lvl <= (int)c->p.cbs_level; lvl++) {
        u64 mul = 1ull << (64 - delta_log - 1);
        for (int i = 0; i <= n; i++) shifted[i] = lwe_in[i] * mul;
        shifted[n] += 1ull << 62;
        u64 alpha = 1ull << (64 - 1 - c->p.cbs_base_log * lvl);
        for (int j = 0; j < N; j++) lut[j] = (u64)0 - alpha;
        bootstrap_one(c, shifted.data(), lut.data(), bs.data());
        bs[big] += alpha;
        for (int r = 0; r <= k; r++) orc_pfks(c, r, bs.data(), ggsw_out + ((size_t)(lvl - 1) * (k + 1) + r) * gsz);
    }
}

// ------------------------------------------------------------------------------------------------
// SURVEY §9.4(7) vertical_packing — tfhe-rs wop_pbs::vertical_packing = cmux tree over the MSB
// GGSWs + blind_rotate_assign over the rest (LSB last, monomial degrees 1,2,4,..) + sample extract.
// Call site many_wopbs.rs:267-279.
// ------------------------------------------------------------------------------------------------
static void vertical_packing_f(const orc_ctx *c, const u64 *lut, int npoly, const double *ggsw_f, int nggsw, u64 *lwe_out) {
    const int k = c->k, N = c->N, lvl = c->p.cbs_level, bl = c->p.cbs_base_log;
    const size_t gsz = (size_t)(k + 1) * N;
    const size_t fsz = (size_t)lvl * (k + 1) * (k + 1) * c->M * 2;
    int log_lut = 0; while ((2 << log_lut) <= npoly) log_lut++;
    int tree = (log_lut > nggsw) ? 0 : log_lut;
    // cmux tree: leaves are trivial GLWEs of the LUT polynomials; GGSW[tree-1] (least significant
    // tree bit) muxes neighbouring leaves, GGSW[0] (MSB) is the root.
    int nleaf = 1 << tree;
    std::vector<u64> cur((size_t)nleaf * gsz, 0), diff(gsz);
    for (int q = 0; q < nleaf; q++) memcpy(&cur[(size_t)q * gsz + (size_t)k * N], lut + (size_t)q * N, (size_t)N * 8);
    for (int layer = tree - 1, cnt = nleaf; layer >= 0; layer--, cnt >>= 1) {
        const double *g = ggsw_f + (size_t)layer * fsz;
        for (int q = 0; q < cnt / 2; q++) {
            u64 *c0 = &cur[(size_t)(2 * q) * gsz], *c1 = &cur[(size_t)(2 * q + 1) * gsz];
            for (size_t t = 0; t < gsz; t++) diff[t] = c1[t] - c0[t];
            external_product_add(c, g, bl, lvl, diff.data(), c0);  // c0 += ggsw (x) (c1 - c0)
            if (q) memcpy(&cur[(size_t)q * gsz], c0, gsz * 8);
        }
    }
    // blind rotation with the remaining GGSWs, last (LSB) first, degrees 1,2,4,...
    std::vector<u64> acc(cur.begin(), cur.begin() + gsz), ct1(gsz);
    int deg = 1;
    for (int g = nggsw - 1; g >= tree; g--, deg <<= 1) {
        for (int r = 0; r <= k; r++) {
            poly_div_monomial(&acc[(size_t)r * N], deg % (2 * N), N, &ct1[(size_t)r * N]);
            for (int j = 0; j < N; j++) ct1[(size_t)r * N + j] -= acc[(size_t)r * N + j];
        }
        external_product_add(c, ggsw_f + (size_t)g * fsz, bl, lvl, ct1.data(), acc.data());
    }
    for (int r = 0; r < k; r++) {
        const u64 *a = &acc[(size_t)r * N];
        lwe_out[(size_t)r * N] = a[0];
        for (int j = 1; j < N; j++) lwe_out[(size_t)r * N + j] = (u64)0 - a[N - j];
    }
    lwe_out[(size_t)k * N] = acc[(size_t)k * N];
}
extern "C" void orc_vertical_packing(const orc_ctx *c, const uint64_t *lut, int npoly, const uint64_t *ggsw_std, int nggsw, uint64_t *lwe_out) {
    const size_t ssz = (size_t)c->p.cbs_level * (c->k + 1) * (c->k + 1) * c->N;
    const size_t fsz = (size_t)c->p.cbs_level * (c->k + 1) * (c->k + 1) * c->M * 2;
    std::vector<double> f((size_t)nggsw * fsz);
    for (int g = 0; g < nggsw; g++) ggsw_to_fourier(c, ggsw_std + (size_t)g * ssz, c->p.cbs_level, &f[(size_t)g * fsz]);
    vertical_packing_f(c, lut, npoly, f.data(), nggsw, lwe_out);
}

// ------------------------------------------------------------------------------------------------
// gen_lut — gen_lut.rs:9-42, restated line for line.
// ------------------------------------------------------------------------------------------------
static int ilog2(u64 v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; }
extern "C" int orc_lut_size(const orc_params *p, int nb_block) {
    int log_basis = ilog2(p->message_modulus) + ilog2(p->carry_modulus);
    int sz = 1 << (nb_block * log_basis);
    return sz < (int)p->poly_size ? (int)p->poly_size : sz;
}
extern "C" void orc_gen_lut(const orc_params *p, int nb_block, const uint64_t *table, uint64_t *lut) {
    const u64 log_msg = ilog2(p->message_modulus), log_carry = ilog2(p->carry_modulus);
    const u64 log_basis = log_msg + log_carry, delta = 64 - log_basis;
    const int lut_size = orc_lut_size(p, nb_block);
    const u64 nvals = 1ull << (nb_block * log_basis);
    for (int index = 0; index < lut_size; index++) {
        u64 value = 0, tmp_index = index;
        for (u64 i = 0; i < (u64)nb_block; i++) {
            u64 tmp = tmp_index % (1ull << log_basis);
            tmp_index >>= log_basis;
            value += tmp << (log_msg * i);
        }
        u64 fv = table[value % nvals];
        for (int b = 0; b < nb_block; b++) {
            u64 masked = (fv >> (log_msg * b)) % (1ull << log_msg);
            lut[(size_t)b * lut_size + index] = masked << delta;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// many_wopbs_without_padding — many_wopbs.rs:31-116 (custom_extract_bits :161-202, CBS loop
// :252-264, VP loop :267-279).  Single-LUT wopbs_without_padding (sbox.rs:61, server.rs:150) is
// the L = 1 case of the same chain.
// ------------------------------------------------------------------------------------------------
extern "C" void orc_many_wopbs(const orc_ctx *c, const uint64_t *ct_in, int nblocks, const uint64_t *luts, int L, uint64_t *out) {
    const int n = c->n, big = c->big, N = c->N, k = c->k;
    const orc_params &p = c->p;
    const int block_mod = p.message_modulus * p.carry_modulus;
    const int bits_per_block = ilog2(block_mod);
    const int total_bits = nblocks * bits_per_block;
    const int delta_log = ilog2((1ull << 63) / (block_mod / 2));
    std::vector<u64> extracted((size_t)total_bits * (n + 1));
    // many_wopbs.rs:179: blocks iterated in reverse => list index 0 = MSB of the last block
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < nblocks; bi++) {
        int blk = nblocks - 1 - bi;
        orc_extract_bits(c, ct_in + (size_t)blk * (big + 1), delta_log, bits_per_block, &extracted[(size_t)bi * bits_per_block * (n + 1)]);
    }
    const size_t ssz = (size_t)p.cbs_level * (k + 1) * (k + 1) * N;
    const size_t fsz = (size_t)p.cbs_level * (k + 1) * (k + 1) * c->M * 2;
    std::vector<double> ggsw_f((size_t)total_bits * fsz);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < total_bits; b++) {
        std::vector<u64> ggsw(ssz);
        orc_circuit_bootstrap_boolean(c, &extracted[(size_t)b * (n + 1)], 63, ggsw.data());  // many_wopbs.rs:257
        ggsw_to_fourier(c, ggsw.data(), p.cbs_level, &ggsw_f[(size_t)b * fsz]);              // many_wopbs.rs:263
    }
    const int lut_size = orc_lut_size(&p, nblocks);
    const int small = lut_size / N;  // polynomials per output (many_wopbs.rs:268-270)
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int i = 0; i < L; i++)
        for (int o = 0; o < nblocks; o++)
            vertical_packing_f(c, luts + ((size_t)i * nblocks + o) * lut_size, small, ggsw_f.data(), total_bits,
                               out + ((size_t)i * nblocks + o) * (big + 1));
}

// ------------------------------------------------------------------------------------------------
// tables (table.rs:2-37 = FIPS-197 Fig. 7 / Fig. 14, generated here from the field arithmetic) and
// GF(2^8) constant multiplications (sbox.rs:20-42)
// ------------------------------------------------------------------------------------------------
static uint8_t g_sbox[256], g_inv_sbox[256];
static bool g_tables_ready = false;
static uint8_t xtime(uint8_t x) { return (uint8_t)((x << 1) ^ ((x & 0x80) ? 0x1B : 0)); }
static uint8_t gmul(uint8_t a, uint8_t b) { uint8_t r = 0; while (b) { if (b & 1) r ^= a; a = xtime(a); b >>= 1; } return r; }
static void init_tables() {
    if (g_tables_ready) return;
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        if (x) for (int y = 1; y < 256; y++) if (gmul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        uint8_t s = inv, r = inv;
        for (int i = 0; i < 4; i++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        s ^= 0x63;
        g_sbox[x] = s; g_inv_sbox[s] = (uint8_t)x;
    }
    g_tables_ready = true;
}
extern "C" const uint8_t *orc_sbox_table(int inv) { init_tables(); return inv ? g_inv_sbox : g_sbox; }
extern "C" uint8_t orc_gf_mul(uint8_t x, int m) { return gmul(x, (uint8_t)m); }

static void make_lut(const orc_ctx *c, int nb, const std::vector<u64> &table, u64 *dst) { orc_gen_lut(&c->p, nb, table.data(), dst); }

// sbox.rs:46-63
extern "C" void orc_sbox(const orc_ctx *c, uint64_t *byte, int inv) {
    init_tables();
    const int lsz = orc_lut_size(&c->p, 8);
    std::vector<u64> tab(256), lut((size_t)8 * lsz), out((size_t)8 * (c->big + 1));
    for (int x = 0; x < 256; x++) tab[x] = inv ? g_inv_sbox[x] : g_sbox[x];
    make_lut(c, 8, tab, lut.data());
    orc_many_wopbs(c, byte, 8, lut.data(), 1, out.data());
    memcpy(byte, out.data(), out.size() * 8);
}
// sbox.rs:68-97: enc {S, 2S, 3S}; dec {9x, 11x, 13x, 14x}
extern "C" void orc_many_sbox(const orc_ctx *c, const uint64_t *byte_in, int inv, uint64_t *out) {
    init_tables();
    const int lsz = orc_lut_size(&c->p, 8);
    const int L = inv ? 4 : 3;
    std::vector<u64> tab(256), luts((size_t)L * 8 * lsz);
    static const int dec_m[4] = {9, 11, 13, 14}, enc_m[3] = {1, 2, 3};
    for (int i = 0; i < L; i++) {
        for (int x = 0; x < 256; x++) tab[x] = inv ? gmul((uint8_t)x, dec_m[i]) : gmul(g_sbox[x], enc_m[i]);
        make_lut(c, 8, tab, &luts[(size_t)i * 8 * lsz]);
    }
    orc_many_wopbs(c, byte_in, 8, luts.data(), L, out);
}

// ------------------------------------------------------------------------------------------------
// linear layers — server.rs:278-282, mix_columns.rs:4-78, inv_mix_columns.rs:4-58,
// shift_rows.rs:5-21, inv_shift_rows.rs:5-21.  unchecked_add = wrapping u64 add of all kN+1 words.
// ------------------------------------------------------------------------------------------------
static inline size_t bytesz(const orc_ctx *c) { return (size_t)8 * (c->big + 1); }
static void byte_add(const orc_ctx *c, u64 *dst, const u64 *a, const u64 *b) { size_t s = bytesz(c); for (size_t i = 0; i < s; i++) dst[i] = a[i] + b[i]; }
extern "C" void orc_add_round_key(const orc_ctx *c, uint64_t *state, const uint64_t *rk) {
    size_t s = 16 * bytesz(c);
    for (size_t i = 0; i < s; i++) state[i] += rk[i];
}
static void swap_bytes(const orc_ctx *c, u64 *base, int stride_bytes, int a, int b) {  // stride in encrypted-byte units
    size_t s = bytesz(c) * stride_bytes;
    std::vector<u64> t(base + a * s, base + (a + 1) * s);
    memcpy(base + a * s, base + b * s, s * 8);
    memcpy(base + b * s, t.data(), s * 8);
}
static void shift_rows_generic(const orc_ctx *c, u64 *st, int stride, int inverse) {
    if (!inverse) {  // shift_rows.rs:9-20 / mix_columns.rs:10-21
        swap_bytes(c, st, stride, 1, 5); swap_bytes(c, st, stride, 5, 9); swap_bytes(c, st, stride, 9, 13);
        swap_bytes(c, st, stride, 2, 10); swap_bytes(c, st, stride, 6, 14);
        swap_bytes(c, st, stride, 3, 15); swap_bytes(c, st, stride, 15, 11); swap_bytes(c, st, stride, 11, 7);
    } else {  // inv_shift_rows.rs:9-20
        swap_bytes(c, st, stride, 13, 9); swap_bytes(c, st, stride, 9, 5); swap_bytes(c, st, stride, 5, 1);
        swap_bytes(c, st, stride, 14, 6); swap_bytes(c, st, stride, 10, 2);
        swap_bytes(c, st, stride, 3, 7); swap_bytes(c, st, stride, 7, 11); swap_bytes(c, st, stride, 11, 15);
    }
}
extern "C" void orc_shift_rows(const orc_ctx *c, uint64_t *state, int inverse) { shift_rows_generic(c, state, 1, inverse); }
extern "C" void orc_mix_columns(const orc_ctx *c, const uint64_t *mul_in, uint64_t *state_out) {
    const size_t bs = bytesz(c);
    std::vector<u64> m(mul_in, mul_in + 16 * 3 * bs), t(bs);
    shift_rows_generic(c, m.data(), 3, 0);
    auto S = [&](int byte, int which) { return &m[((size_t)byte * 3 + which) * bs]; };
    for (int col = 0; col < 4; col++) {
        int b = col * 4;
        u64 *o = state_out + (size_t)b * bs;
        // mix_columns.rs:36-75 ([0]=S, [1]=2S, [2]=3S)
        byte_add(c, t.data(), S(b + 2, 0), S(b + 3, 0)); byte_add(c, t.data(), S(b + 1, 2), t.data()); byte_add(c, o, S(b, 1), t.data());
        byte_add(c, t.data(), S(b + 2, 2), S(b + 3, 0)); byte_add(c, t.data(), S(b + 1, 1), t.data()); byte_add(c, o + bs, S(b, 0), t.data());
        byte_add(c, t.data(), S(b + 2, 1), S(b + 3, 2)); byte_add(c, t.data(), S(b + 1, 0), t.data()); byte_add(c, o + 2 * bs, S(b, 0), t.data());
        byte_add(c, t.data(), S(b + 2, 0), S(b + 3, 1)); byte_add(c, t.data(), S(b + 1, 0), t.data()); byte_add(c, o + 3 * bs, S(b, 2), t.data());
    }
}
extern "C" void orc_inv_mix_columns(const orc_ctx *c, const uint64_t *m, uint64_t *state_out) {
    const size_t bs = bytesz(c);
    std::vector<u64> t(bs);
    auto S = [&](int byte, int which) { return m + ((size_t)byte * 4 + which) * bs; };
    for (int col = 0; col < 4; col++) {
        int b = col * 4;
        u64 *o = state_out + (size_t)b * bs;
        // inv_mix_columns.rs:17-55 ([0]=9x, [1]=11x, [2]=13x, [3]=14x)
        byte_add(c, t.data(), S(b + 2, 2), S(b + 3, 0)); byte_add(c, t.data(), S(b + 1, 1), t.data()); byte_add(c, o, S(b, 3), t.data());
        byte_add(c, t.data(), S(b + 2, 1), S(b + 3, 2)); byte_add(c, t.data(), S(b + 1, 3), t.data()); byte_add(c, o + bs, S(b, 0), t.data());
        byte_add(c, t.data(), S(b + 2, 3), S(b + 3, 1)); byte_add(c, t.data(), S(b + 1, 0), t.data()); byte_add(c, o + 2 * bs, S(b, 2), t.data());
        byte_add(c, t.data(), S(b + 2, 0), S(b + 3, 3)); byte_add(c, t.data(), S(b + 1, 2), t.data()); byte_add(c, o + 3 * bs, S(b, 1), t.data());
    }
}

// ------------------------------------------------------------------------------------------------
// Server::aes_encrypt — server.rs:39-64
// ------------------------------------------------------------------------------------------------
extern "C" void orc_aes_encrypt(const orc_ctx *c, const uint64_t *rk, uint64_t *state) {
    const size_t bs = bytesz(c);
    std::vector<u64> mul((size_t)16 * 3 * bs), ns((size_t)16 * bs);
    orc_add_round_key(c, state, rk);
    for (int round = 1; round < 10; round++) {
        for (int b = 0; b < 16; b++) orc_many_sbox(c, state + (size_t)b * bs, 0, &mul[(size_t)b * 3 * bs]);
        orc_mix_columns(c, mul.data(), ns.data());
        orc_add_round_key(c, ns.data(), rk + (size_t)round * 16 * bs);
        memcpy(state, ns.data(), ns.size() * 8);
    }
    for (int b = 0; b < 16; b++) orc_sbox(c, state + (size_t)b * bs, 0);
    orc_shift_rows(c, state, 0);
    orc_add_round_key(c, state, rk + (size_t)10 * 16 * bs);
}
// Server::aes_decrypt — server.rs:67-105
extern "C" void orc_aes_decrypt(const orc_ctx *c, const uint64_t *rk, uint64_t *state) {
    const size_t bs = bytesz(c);
    std::vector<u64> mul((size_t)16 * 4 * bs);
    orc_add_round_key(c, state, rk + (size_t)10 * 16 * bs);
    for (int round = 10; round >= 2; round--) {
        orc_shift_rows(c, state, 1);
        for (int b = 0; b < 16; b++) orc_sbox(c, state + (size_t)b * bs, 1);
        orc_add_round_key(c, state, rk + (size_t)(round - 1) * 16 * bs);
        for (int b = 0; b < 16; b++) orc_many_sbox(c, state + (size_t)b * bs, 1, &mul[(size_t)b * 4 * bs]);
        orc_inv_mix_columns(c, mul.data(), state);
    }
    orc_shift_rows(c, state, 1);
    for (int b = 0; b < 16; b++) orc_sbox(c, state + (size_t)b * bs, 1);
    orc_add_round_key(c, state, rk);
}
// Server::aes_key_expansion — server.rs:107-167 (+ key_expansion_utils.rs:10-28)
extern "C" void orc_aes_key_expansion(const orc_ctx *c, const uint64_t *key_ct, const uint64_t *rcon_ct, uint64_t *round_keys) {
    static const uint8_t RCON[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36};
    const size_t bs = bytesz(c);
    const int lsz = orc_lut_size(&c->p, 8);
    std::vector<u64> ident(256), idlut((size_t)8 * lsz);
    for (int x = 0; x < 256; x++) ident[x] = x;
    make_lut(c, 8, ident, idlut.data());                       // server.rs:118-119
    u64 *w = round_keys;                                       // w[i] = word i = 4 bytes, flat == round-key layout
    memcpy(w, key_ct, (size_t)16 * bs * 8);                    // server.rs:122-128
    std::vector<u64> temp((size_t)4 * bs), rc(bs), sum(bs), ref(bs);
    for (int i = 4; i < 44; i++) {
        memcpy(temp.data(), w + (size_t)(i - 1) * 4 * bs, (size_t)4 * bs * 8);
        if (i % 4 == 0) {
            std::vector<u64> rot((size_t)4 * bs);              // fhe_rot_word
            for (int j = 0; j < 4; j++) memcpy(&rot[(size_t)j * bs], &temp[(size_t)((j + 1) % 4) * bs], bs * 8);
            temp = rot;
            for (int j = 0; j < 4; j++) orc_sbox(c, &temp[(size_t)j * bs], 0);  // fhe_sub_word
            if (rcon_ct) memcpy(rc.data(), rcon_ct + (size_t)(i / 4 - 1) * bs, bs * 8);
            else orc_trivial_bytes(c, &RCON[i / 4 - 1], 1, rc.data());
            byte_add(c, temp.data(), temp.data(), rc.data());  // server.rs:143
        }
        for (int j = 0; j < 4; j++) {
            byte_add(c, sum.data(), w + ((size_t)(i - 4) * 4 + j) * bs, &temp[(size_t)j * bs]);  // server.rs:148
            orc_many_wopbs(c, sum.data(), 8, idlut.data(), 1, ref.data());                        // server.rs:150
            memcpy(w + ((size_t)i * 4 + j) * bs, ref.data(), bs * 8);
        }
    }
}
// Server::add_scalar — server.rs:172-275.  faithful=1 keeps the whole counter in the low-byte
// LUTs (server.rs:181-182), which is wrong for i >= 256; faithful=0 uses i & 0xFF (FIPS / aes crate).
extern "C" void orc_add_scalar(const orc_ctx *c, uint64_t *state, uint64_t lo, uint64_t hi, int faithful) {
    const size_t bs = bytesz(c), lw = (size_t)c->big + 1;
    uint8_t ib[16];
    for (int j = 0; j < 16; j++) { int sh = 8 * (15 - j); ib[j] = (uint8_t)((sh >= 64 ? hi >> (sh - 64) : lo >> sh) & 0xFF); }
    const u64 addend = faithful ? lo : (lo & 0xFF);
    std::vector<u64> ns((size_t)16 * bs);
    {
        const int lsz = orc_lut_size(&c->p, 8);
        std::vector<u64> tf(256), tg(256), luts((size_t)2 * 8 * lsz), res((size_t)2 * bs);
        for (u64 x = 0; x < 256; x++) { tf[x] = (x + addend) % 256; tg[x] = (x + addend > 255) ? 1 : 0; }
        make_lut(c, 8, tf, luts.data()); make_lut(c, 8, tg, &luts[(size_t)8 * lsz]);
        orc_many_wopbs(c, state + (size_t)15 * bs, 8, luts.data(), 2, res.data());
        memcpy(&ns[(size_t)15 * bs], res.data(), bs * 8);
        std::vector<u64> carry(res.begin() + bs, res.begin() + bs + lw);  // block 0 of the carry result
        const int lsz9 = orc_lut_size(&c->p, 9);
        std::vector<u64> t9f(512), t9g(512), luts9((size_t)2 * 9 * lsz9), in9((size_t)9 * lw), res9((size_t)2 * 9 * lw);
        for (int index = 14; index >= 0; index--) {
            memcpy(in9.data(), state + (size_t)index * bs, bs * 8);
            memcpy(&in9[(size_t)8 * lw], carry.data(), lw * 8);
            for (u64 x = 0; x < 512; x++) {
                u64 s = (x & 0xFF) + ((x >> 8) & 1) + ib[index];
                t9f[x] = s % 256; t9g[x] = s > 255 ? 1 : 0;
            }
            make_lut(c, 9, t9f, luts9.data()); make_lut(c, 9, t9g, &luts9[(size_t)9 * lsz9]);
            orc_many_wopbs(c, in9.data(), 9, luts9.data(), 2, res9.data());
            memcpy(&ns[(size_t)index * bs], res9.data(), bs * 8);               // blocks 0..7 of the sum
            memcpy(carry.data(), &res9[(size_t)9 * lw], lw * 8);               // block 0 of the carry
        }
    }
    memcpy(state, ns.data(), ns.size() * 8);
}

// ------------------------------------------------------------------------------------------------
// clear AES-128 (FIPS-197) — the expected values the reference takes from the `aes` crate
// (client.rs:163-171)
// ------------------------------------------------------------------------------------------------
extern "C" void orc_clear_aes_key_expansion(const uint8_t key[16], uint8_t rk[176]) {
    init_tables();
    static const uint8_t RCON[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36};
    memcpy(rk, key, 16);
    for (int i = 4; i < 44; i++) {
        uint8_t t[4]; memcpy(t, rk + 4 * (i - 1), 4);
        if (i % 4 == 0) { uint8_t u = t[0]; t[0] = g_sbox[t[1]] ^ RCON[i / 4 - 1]; t[1] = g_sbox[t[2]]; t[2] = g_sbox[t[3]]; t[3] = g_sbox[u]; }
        for (int j = 0; j < 4; j++) rk[4 * i + j] = rk[4 * (i - 4) + j] ^ t[j];
    }
}
extern "C" void orc_clear_aes_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]) {
    uint8_t rk[176], s[16], t[16];
    orc_clear_aes_key_expansion(key, rk);
    for (int i = 0; i < 16; i++) s[i] = in[i] ^ rk[i];
    for (int r = 1; r <= 10; r++) {
        for (int i = 0; i < 16; i++) s[i] = g_sbox[s[i]];
        for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) t[4 * col + row] = s[4 * ((col + row) % 4) + row];
        if (r < 10) for (int col = 0; col < 4; col++) {
            uint8_t *a = t + 4 * col, b0 = a[0], b1 = a[1], b2 = a[2], b3 = a[3];
            a[0] = gmul(b0, 2) ^ gmul(b1, 3) ^ b2 ^ b3; a[1] = b0 ^ gmul(b1, 2) ^ gmul(b2, 3) ^ b3;
            a[2] = b0 ^ b1 ^ gmul(b2, 2) ^ gmul(b3, 3); a[3] = gmul(b0, 3) ^ b1 ^ b2 ^ gmul(b3, 2);
        }
        for (int i = 0; i < 16; i++) s[i] = t[i] ^ rk[16 * r + i];
    }
    memcpy(out, s, 16);
}
extern "C" void orc_clear_aes_decrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]) {
    uint8_t rk[176], s[16], t[16];
    orc_clear_aes_key_expansion(key, rk);
    for (int i = 0; i < 16; i++) s[i] = in[i] ^ rk[160 + i];
    for (int r = 9; r >= 0; r--) {
        for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) t[4 * ((col + row) % 4) + row] = s[4 * col + row];
        for (int i = 0; i < 16; i++) s[i] = g_inv_sbox[t[i]] ^ rk[16 * r + i];
        if (r > 0) for (int col = 0; col < 4; col++) {
            uint8_t *a = s + 4 * col, b0 = a[0], b1 = a[1], b2 = a[2], b3 = a[3];
            a[0] = gmul(b0, 14) ^ gmul(b1, 11) ^ gmul(b2, 13) ^ gmul(b3, 9); a[1] = gmul(b0, 9) ^ gmul(b1, 14) ^ gmul(b2, 11) ^ gmul(b3, 13);
            a[2] = gmul(b0, 13) ^ gmul(b1, 9) ^ gmul(b2, 14) ^ gmul(b3, 11); a[3] = gmul(b0, 11) ^ gmul(b1, 13) ^ gmul(b2, 9) ^ gmul(b3, 14);
        }
    }
    memcpy(out, s, 16);
}
