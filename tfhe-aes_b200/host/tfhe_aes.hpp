// tfhe_aes.hpp — C++ host-side mirror of the reference's Rust interface for the hot path, on top of the
// C ABI (include/tfhe_aes_b200.h).  Same names, argument meaning and error behaviour as the reference:
//   Server::{new -> constructor, aes_key_expansion, aes_encrypt, aes_decrypt, add_scalar}   (server.rs:24-282)
//   sbox::{gen_lut, many_wopbs_without_padding, sbox, many_sbox, mul2..mul14}                 (sbox/*.rs)
//   Client::{new -> constructor, client_encrypt, client_decrypt_and_verify}                   (client.rs:70-175)
// Containers are the flat u64 layouts of tfhe-rs: a radix byte is 8 LWEs of lw words (block j = bit j),
// a state is 16 bytes, round keys are 11 states.  The reference panics on error; this mirror throws.
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/tfhe_aes_b200.h"

namespace tfhe_aes {

using u64 = uint64_t;
using u128 = unsigned __int128;
using RadixCiphertext = std::vector<u64>;              // [8][lw]   BaseRadixCiphertext<Ciphertext>
using State = std::vector<u64>;                         // [16][8][lw]
using RoundKeys = std::vector<u64>;                     // [11][16][8][lw]
using IntegerWopbsLUT = std::vector<u64>;               // [nb_block][lut_size]

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

// PARAM_OPT (client.rs:31-57)
inline tfa_params param_opt() { tfa_params p; tfa_param_opt(&p); return p; }

class Engine {  // owns the GPU context = (public_key, sks, wopbs_key) of Server::new (server.rs:32)
public:
    explicit Engine(const tfa_params &p, int device = 0) : params(p) {
        if (tfa_ctx_create(&p, device, nullptr, &ctx) != TFA_OK) throw Error(tfa_last_error(nullptr));
        lw = p.glwe_dim * p.poly_size + 1;
    }
    ~Engine() { tfa_ctx_destroy(ctx); }
    Engine(const Engine &) = delete;
    void check(int rc) const { if (rc != TFA_OK) throw Error(tfa_last_error(ctx)); }
    void load_keys(const u64 *bsk, const u64 *ksk, const u64 *pfpksk) { check(tfa_ctx_load_keys(ctx, bsk, ksk, pfpksk)); }
    tfa_ctx *ctx = nullptr;
    tfa_params params;
    size_t lw = 0;
    size_t byte_words() const { return 8 * lw; }
    size_t state_words() const { return 16 * byte_words(); }
};

namespace sbox {
// sbox.rs:20-42
inline uint8_t mul2(uint8_t x) { return (uint8_t)((x << 1) ^ ((x & 0x80) ? 0x1B : 0)); }
inline uint8_t mul3(uint8_t x) { return mul2(x) ^ x; }
inline uint8_t mul9(uint8_t x) { return mul2(mul2(mul2(x))) ^ x; }
inline uint8_t mul11(uint8_t x) { return mul2(mul2(mul2(x)) ^ x) ^ x; }
inline uint8_t mul13(uint8_t x) { return mul2(mul2(mul2(x) ^ x)) ^ x; }
inline uint8_t mul14(uint8_t x) { return mul2(mul2(mul2(x) ^ x) ^ x); }

// gen_lut.rs:9
inline IntegerWopbsLUT gen_lut(const tfa_params &p, int nb_block, const std::function<u64(u64)> &f) {
    int log_basis = 0;
    for (u64 m = (u64)p.message_modulus * p.carry_modulus; m > 1; m >>= 1) log_basis++;
    std::vector<u64> table((size_t)1 << (nb_block * log_basis));
    for (size_t v = 0; v < table.size(); v++) table[v] = f(v);
    const int size = tfa_lut_size(&p, nb_block);
    if (size < 0) throw Error("gen_lut: bad nb_block");
    IntegerWopbsLUT lut((size_t)nb_block * size);
    if (tfa_gen_lut(&p, nb_block, table.data(), lut.data()) != TFA_OK) throw Error("gen_lut failed");
    return lut;
}
// many_wopbs.rs:31
inline std::vector<RadixCiphertext> many_wopbs_without_padding(Engine &e, const RadixCiphertext &ct_in, const std::vector<IntegerWopbsLUT> &luts) {
    const int nblocks = (int)(ct_in.size() / e.lw), L = (int)luts.size();
    std::vector<u64> flat;
    for (auto &l : luts) flat.insert(flat.end(), l.begin(), l.end());
    std::vector<u64> out((size_t)L * nblocks * e.lw);
    e.check(tfa_many_wopbs(e.ctx, ct_in.data(), 1, nblocks, flat.data(), L, out.data()));
    std::vector<RadixCiphertext> res;
    for (int i = 0; i < L; i++) res.emplace_back(out.begin() + (size_t)i * nblocks * e.lw, out.begin() + (size_t)(i + 1) * nblocks * e.lw);
    return res;
}
// sbox.rs:46 (in place) and sbox.rs:68
inline void sbox(Engine &e, RadixCiphertext &ct_in, bool inv) { e.check(tfa_sbox(e.ctx, ct_in.data(), 1, inv)); }
inline std::vector<RadixCiphertext> many_sbox(Engine &e, const RadixCiphertext &ct_in, bool inv) {
    const int L = inv ? 4 : 3;
    std::vector<u64> out((size_t)L * e.byte_words());
    e.check(tfa_many_sbox(e.ctx, ct_in.data(), 1, inv, out.data()));
    std::vector<RadixCiphertext> res;
    for (int i = 0; i < L; i++) res.emplace_back(out.begin() + (size_t)i * e.byte_words(), out.begin() + (size_t)(i + 1) * e.byte_words());
    return res;
}
}  // namespace sbox

class Server {  // server.rs:24-35
public:
    explicit Server(Engine &engine) : e(engine) {}
    // server.rs:107
    RoundKeys aes_key_expansion(const State &key) const {
        RoundKeys rk(11 * e.state_words());
        e.check(tfa_aes_key_expansion(e.ctx, key.data(), nullptr, rk.data()));
        return rk;
    }
    // server.rs:39 / :67 — one state, in place
    void aes_encrypt(const RoundKeys &rk, State &state) const { e.check(tfa_aes_encrypt(e.ctx, rk.data(), state.data(), (int)(state.size() / e.state_words()))); }
    void aes_decrypt(const RoundKeys &rk, State &state) const { e.check(tfa_aes_decrypt(e.ctx, rk.data(), state.data(), (int)(state.size() / e.state_words()))); }
    void aes_encryption(const RoundKeys &rk, State &s) const { aes_encrypt(rk, s); }   // README.md:58-59
    void aes_decryption(const RoundKeys &rk, State &s) const { aes_decrypt(rk, s); }
    // server.rs:172 — state += i (low-byte LUT uses i & 0xFF: correct for i >= 256, unlike server.rs:181-182)
    void add_scalar(State &state, u128 i) const {
        const u64 ctr[2] = {(u64)i, (u64)(i >> 64)};
        e.check(tfa_add_scalar(e.ctx, state.data(), ctr, 1));
    }
    // main.rs:55-64 — the rayon CTR loop as one batched call: out[b] = AES(iv + first + b)
    std::vector<u64> aes_ctr(const RoundKeys &rk, const State &iv, int number_of_outputs, u128 first = 0) const {
        std::vector<u64> out((size_t)number_of_outputs * e.state_words());
        e.check(tfa_aes_ctr(e.ctx, rk.data(), iv.data(), (u64)first, (u64)(first >> 64), number_of_outputs, out.data()));
        return out;
    }
private:
    Engine &e;
};

class Client {  // client.rs:59-175 (trusted side: keygen, encrypt, decrypt + verify) on the GPU harness
public:
    // seed 0 = keys and encryption randomness from the operating system's entropy; a non-zero seed is reproducible and INSECURE (tests)
    Client(Engine &engine, size_t number_of_outputs, u128 iv, u128 key, u64 seed = 0) : e(engine), n_out(number_of_outputs), iv_(iv), key_(key), seed_(seed) {
        e.check(tfa_client_keygen(e.ctx, seed));    // gen_keys_radix + new_wopbs_key_only_for_wopbs (client.rs:106-107)
    }
    static void to_bytes(u128 v, uint8_t out[16]) { for (int i = 0; i < 16; i++) out[i] = (uint8_t)(v >> (8 * (15 - i))); }  // MSB byte first (client.rs:126-129)
    // client.rs:123 — returns (encrypted_iv, encrypted_key)
    std::pair<State, State> client_encrypt() const {
        uint8_t kb[16], ib[16];
        to_bytes(key_, kb); to_bytes(iv_, ib);
        State k(e.state_words()), i(e.state_words());
        e.check(tfa_client_encrypt_bytes(e.ctx, kb, 16, seed_ ? seed_ + 11 : 0, k.data()));
        e.check(tfa_client_encrypt_bytes(e.ctx, ib, 16, seed_ ? seed_ + 12 : 0, i.data()));
        return {i, k};
    }
    std::vector<uint8_t> decrypt(const std::vector<u64> &states) const {
        const int nbytes = (int)(states.size() / e.byte_words());
        std::vector<uint8_t> out(nbytes);
        e.check(tfa_client_decrypt_bytes(e.ctx, states.data(), nbytes, out.data()));
        return out;
    }
    size_t number_of_outputs() const { return n_out; }
    u128 iv() const { return iv_; }
    u128 key() const { return key_; }
private:
    Engine &e;
    size_t n_out;
    u128 iv_, key_;
    u64 seed_;
};

// FIPS-197 AES-128 in the clear (the `aes` crate of client.rs:163-171), for client_decrypt_and_verify
namespace clear {
inline const uint8_t *sbox_table() {
    static uint8_t sb[256];
    static bool init = false;
    if (!init) {
        auto gm = [](uint8_t a, uint8_t b) { uint8_t r = 0; while (b) { if (b & 1) r ^= a; a = sbox::mul2(a); b >>= 1; } return r; };
        for (int x = 0; x < 256; x++) {
            uint8_t inv = 0;
            if (x) for (int y = 1; y < 256; y++) if (gm((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
            uint8_t s = inv, r = inv;
            for (int i = 0; i < 4; i++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
            sb[x] = s ^ 0x63;
        }
        init = true;
    }
    return sb;
}
inline void aes128_encrypt(const uint8_t key[16], const uint8_t in[16], uint8_t out[16]) {
    static const uint8_t RCON[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36};
    const uint8_t *sb = sbox_table();
    uint8_t rk[176];
    for (int i = 0; i < 16; i++) rk[i] = key[i];
    for (int i = 4; i < 44; i++) {
        uint8_t t[4] = {rk[4 * i - 4], rk[4 * i - 3], rk[4 * i - 2], rk[4 * i - 1]};
        if (i % 4 == 0) { uint8_t u = t[0]; t[0] = sb[t[1]] ^ RCON[i / 4 - 1]; t[1] = sb[t[2]]; t[2] = sb[t[3]]; t[3] = sb[u]; }
        for (int j = 0; j < 4; j++) rk[4 * i + j] = rk[4 * (i - 4) + j] ^ t[j];
    }
    uint8_t s[16], t[16];
    for (int i = 0; i < 16; i++) s[i] = in[i] ^ rk[i];
    for (int r = 1; r <= 10; r++) {
        for (int i = 0; i < 16; i++) s[i] = sb[s[i]];
        for (int c = 0; c < 4; c++) for (int row = 0; row < 4; row++) t[4 * c + row] = s[4 * ((c + row) % 4) + row];
        if (r < 10) for (int c = 0; c < 4; c++) {
            uint8_t *a = t + 4 * c, b0 = a[0], b1 = a[1], b2 = a[2], b3 = a[3];
            a[0] = sbox::mul2(b0) ^ sbox::mul3(b1) ^ b2 ^ b3; a[1] = b0 ^ sbox::mul2(b1) ^ sbox::mul3(b2) ^ b3;
            a[2] = b0 ^ b1 ^ sbox::mul2(b2) ^ sbox::mul3(b3); a[3] = sbox::mul3(b0) ^ b1 ^ b2 ^ sbox::mul2(b3);
        }
        for (int i = 0; i < 16; i++) s[i] = t[i] ^ rk[16 * r + i];
    }
    for (int i = 0; i < 16; i++) out[i] = s[i];
}
}  // namespace clear

// client.rs:147-175: decrypt every output block and compare with AES-128(key, iv + index); throws on mismatch
inline void client_decrypt_and_verify(const Client &c, const std::vector<u64> &states) {
    const std::vector<uint8_t> dec = c.decrypt(states);
    if (dec.size() != 16 * c.number_of_outputs()) throw Error("client_decrypt_and_verify: wrong number of outputs");
    uint8_t kb[16];
    Client::to_bytes(c.key(), kb);
    for (size_t index = 0; index < c.number_of_outputs(); index++) {
        uint8_t msg[16], exp[16];
        Client::to_bytes(c.iv() + (u128)index, msg);
        clear::aes128_encrypt(kb, msg, exp);
        for (int i = 0; i < 16; i++)
            if (dec[16 * index + i] != exp[i]) throw Error("FHE AES output differs from AES-128 at block " + std::to_string(index));
    }
}
}  // namespace tfhe_aes
