/*
 * tfhe_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY — never part of the product path).
 *
 * A dependency-free restatement of the arithmetic the reference (rostin79s/TFHE-AES) executes
 * through the third-party crate `tfhe 0.11.2` (+ `tfhe-fft 0.7.0`), which is NOT vendored in
 * /root/reference (Cargo.lock:546-566, :580) and cannot be built here (no cargo/rustc, no network).
 *
 * PARITY STATUS:
 *   - plaintext level: PINNED.  Decrypted outputs are checked against FIPS-197 / SP 800-38A known
 *     answers, which is the only check the reference itself holds (client.rs:171,200,214;
 *     main.rs:78-95).
 *   - ciphertext level vs real tfhe-rs: "PARITY UNPINNED".  The reference fixes no LWE/GLWE/GGSW
 *     value anywhere and tfhe-rs keygen is seeded from OS entropy, so no ciphertext-level golden
 *     vector exists.  The algorithms below follow the published tfhe-rs 0.11 algorithms
 *      at the reference's own call sites, cited per function.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
    uint32_t lwe_dim;       /* n   (client.rs:33)  */
    uint32_t glwe_dim;      /* k   (client.rs:34)  */
    uint32_t poly_size;     /* N   (client.rs:35)  */
    uint32_t pbs_base_log;  /* client.rs:42 */
    uint32_t pbs_level;     /* client.rs:43 */
    uint32_t ks_base_log;   /* client.rs:45 */
    uint32_t ks_level;      /* client.rs:44 */
    uint32_t pfks_base_log; /* client.rs:47 */
    uint32_t pfks_level;    /* client.rs:46 */
    uint32_t cbs_base_log;  /* client.rs:52 */
    uint32_t cbs_level;     /* client.rs:51 */
    uint32_t message_modulus; /* client.rs:53 */
    uint32_t carry_modulus;   /* client.rs:54 */
    uint32_t _pad;
    double lwe_std;         /* client.rs:36-38 */
    double glwe_std;        /* client.rs:39-41 */
    double pfks_std;        /* client.rs:48-50 */
} orc_params;

typedef struct orc_ctx orc_ctx;

orc_ctx *orc_create(const orc_params *p);
void orc_destroy(orc_ctx *c);
void orc_set_threads(int nthreads);
int orc_max_threads(void);

/* keygen (client.rs:106-107): secret keys, BSK (standard + Fourier), KSK, 5-key PFPKSK list */
void orc_keygen(orc_ctx *c, uint64_t seed);
const uint64_t *orc_lwe_sk(const orc_ctx *c);   /* [n] binary             */
const uint64_t *orc_glwe_sk(const orc_ctx *c);  /* [k*N] binary           */
const uint64_t *orc_bsk(const orc_ctx *c);      /* [n][l][k+1][(k+1)*N], level 1 first */
const uint64_t *orc_ksk(const orc_ctx *c);      /* [k*N][l_ks][n+1],      level 1 first */
const uint64_t *orc_pfpksk(const orc_ctx *c);   /* [k+1][k*N+1][l_pfks][(k+1)*N], level 1 first */

/* client-side encodings (client.rs:126-138,147-175) */
void orc_seed_encryption(orc_ctx *c, uint64_t seed);
void orc_encrypt_bits(orc_ctx *c, const uint8_t *bits, int count, uint64_t *out); /* big key, bit<<63 */
void orc_encrypt_bytes(orc_ctx *c, const uint8_t *bytes, int count, uint64_t *out); /* [count][8][kN+1], block j = bit j */
void orc_trivial_bytes(const orc_ctx *c, const uint8_t *bytes, int count, uint64_t *out);
void orc_decrypt_bits(const orc_ctx *c, const uint64_t *ct, int count, uint8_t *bits, int64_t *err);
void orc_decrypt_bytes(const orc_ctx *c, const uint64_t *ct, int count, uint8_t *bytes);
void orc_encrypt_lwe_small(orc_ctx *c, const uint64_t *plain, int count, uint64_t *out); /* small key */
void orc_phase_small(const orc_ctx *c, const uint64_t *ct, int count, uint64_t *phase);
void orc_phase_big(const orc_ctx *c, const uint64_t *ct, int count, uint64_t *phase);
void orc_glwe_phase(const orc_ctx *c, const uint64_t *glwe, int count, uint64_t *phase); /* [count][N] */

/* primitives */
void orc_decompose(uint64_t x, int base_log, int level, int64_t *digits /* [level], index 0 = level `level` (first out) */);
void orc_fft_forward_torus(const orc_ctx *c, const uint64_t *poly, double *out /* [N/2][2] */);
void orc_fft_forward_integer(const orc_ctx *c, const int64_t *poly, double *out);
void orc_fft_add_backward_torus(const orc_ctx *c, const double *fourier, uint64_t *poly_inout);
void orc_keyswitch(const orc_ctx *c, const uint64_t *in, int count, uint64_t *out);
/* lut: N coefficients (body of a trivial GLWE) */
void orc_bootstrap(const orc_ctx *c, const uint64_t *in, int count, const uint64_t *lut, uint64_t *out);
void orc_extract_bits(const orc_ctx *c, const uint64_t *in, int delta_log, int nbits, uint64_t *out);
void orc_pfks(const orc_ctx *c, int key_index, const uint64_t *lwe_in, uint64_t *glwe_out);
/* out: standard GGSW [cbs_level][k+1][(k+1)*N] */
void orc_circuit_bootstrap_boolean(const orc_ctx *c, const uint64_t *lwe_in, int delta_log, uint64_t *ggsw_out);
/* ggsw_std: [nggsw][cbs_level][k+1][(k+1)N], index 0 = MSB.  lut: [npoly][N]. */
void orc_vertical_packing(const orc_ctx *c, const uint64_t *lut, int npoly, const uint64_t *ggsw_std,
                          int nggsw, uint64_t *lwe_out);
/* external product / cmux on standard-domain GGSW with arbitrary (base_log, level) */
void orc_external_product_add(const orc_ctx *c, const uint64_t *ggsw_std, int base_log, int level,
                              const uint64_t *glwe_in, uint64_t *glwe_inout);

/* sbox module (gen_lut.rs:9-42, many_wopbs.rs:31-116, sbox.rs:46-97) */
void orc_gen_lut(const orc_params *p, int nb_block, const uint64_t *table /* f(v), v < 2^(nb_block*log_basis) */,
                 uint64_t *lut_out /* [nb_block][lut_size] */);
int orc_lut_size(const orc_params *p, int nb_block);
/* ct_in: [nblocks][kN+1] block 0 = LSB.  luts: [L][nblocks][lut_size].  out: [L][nblocks][kN+1] */
void orc_many_wopbs(const orc_ctx *c, const uint64_t *ct_in, int nblocks, const uint64_t *luts, int L,
                    uint64_t *out);
void orc_sbox(const orc_ctx *c, uint64_t *byte_inout, int inv);
void orc_many_sbox(const orc_ctx *c, const uint64_t *byte_in, int inv, uint64_t *out /* [3 or 4][8][kN+1] */);

/* Server (server.rs:39-282).  state: [16][8][kN+1]; round keys: [11][16][8][kN+1]. */
void orc_aes_key_expansion(const orc_ctx *c, const uint64_t *key_ct, const uint64_t *rcon_ct /* [10][8][kN+1] or NULL => trivial */,
                           uint64_t *round_keys);
void orc_aes_encrypt(const orc_ctx *c, const uint64_t *round_keys, uint64_t *state);
void orc_aes_decrypt(const orc_ctx *c, const uint64_t *round_keys, uint64_t *state);
/* faithful=1 reproduces server.rs:181-182 (whole counter in the low-byte LUT; wrong for i >= 256) */
void orc_add_scalar(const orc_ctx *c, uint64_t *state, uint64_t counter_lo, uint64_t counter_hi, int faithful);
void orc_add_round_key(const orc_ctx *c, uint64_t *state, const uint64_t *rk);
void orc_mix_columns(const orc_ctx *c, const uint64_t *mul_sbox_state /* [16][3][8][kN+1] */, uint64_t *state_out);
void orc_inv_mix_columns(const orc_ctx *c, const uint64_t *mul_state /* [16][4][8][kN+1] */, uint64_t *state_out);
void orc_shift_rows(const orc_ctx *c, uint64_t *state, int inverse);

/* clear AES-128 (FIPS-197) for expected values */
void orc_clear_aes_key_expansion(const uint8_t key[16], uint8_t rk[176]);
void orc_clear_aes_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]);
void orc_clear_aes_decrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]);
const uint8_t *orc_sbox_table(int inv);
uint8_t orc_gf_mul(uint8_t x, int m);

#ifdef __cplusplus
}
#endif
#endif
