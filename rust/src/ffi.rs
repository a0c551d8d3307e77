//! The C ABI of `include/tfhe_aes_b200.h`, one declaration per entry point the shim uses.
use std::ffi::{c_char, c_int, c_void};

/// `tfa_params` (include/tfhe_aes_b200.h) = `WopbsParameters` of the reference (client.rs:31-57).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct TfaParams {
    pub lwe_dim: u32,
    pub glwe_dim: u32,
    pub poly_size: u32,
    pub pbs_base_log: u32,
    pub pbs_level: u32,
    pub ks_base_log: u32,
    pub ks_level: u32,
    pub pfks_base_log: u32,
    pub pfks_level: u32,
    pub cbs_base_log: u32,
    pub cbs_level: u32,
    pub message_modulus: u32,
    pub carry_modulus: u32,
    pub _pad: u32,
    pub lwe_std: f64,
    pub glwe_std: f64,
    pub pfks_std: f64,
}

extern "C" {
    pub fn tfa_ctx_create(p: *const TfaParams, device: c_int, stream: *mut c_void, out: *mut *mut c_void) -> c_int;
    pub fn tfa_ctx_destroy(ctx: *mut c_void);
    pub fn tfa_last_error(ctx: *const c_void) -> *const c_char;
    pub fn tfa_ctx_load_keys(ctx: *mut c_void, bsk: *const u64, ksk: *const u64, pfpksk: *const u64) -> c_int;
    pub fn tfa_aes_key_expansion(ctx: *mut c_void, key: *const u64, rcon: *const u64, rk_out: *mut u64) -> c_int;
    pub fn tfa_aes_encrypt(ctx: *mut c_void, rk: *const u64, states: *mut u64, nblk: c_int) -> c_int;
    pub fn tfa_aes_decrypt(ctx: *mut c_void, rk: *const u64, states: *mut u64, nblk: c_int) -> c_int;
    pub fn tfa_add_scalar(ctx: *mut c_void, states: *mut u64, counters: *const u64, nblk: c_int) -> c_int;
    pub fn tfa_aes_ctr(ctx: *mut c_void, rk: *const u64, iv: *const u64, first_lo: u64, first_hi: u64, nblk: c_int, out: *mut u64) -> c_int;
    pub fn tfa_many_wopbs(ctx: *mut c_void, ct: *const u64, nct: c_int, nblocks: c_int, luts: *const u64, nluts: c_int, out: *mut u64) -> c_int;
    pub fn tfa_sbox(ctx: *mut c_void, bytes: *mut u64, count: c_int, inverse: c_int) -> c_int;
    pub fn tfa_many_sbox(ctx: *mut c_void, bytes: *const u64, count: c_int, inverse: c_int, out: *mut u64) -> c_int;
    pub fn tfa_gen_lut(p: *const TfaParams, nb_block: c_int, table: *const u64, out: *mut u64) -> c_int;
    pub fn tfa_lut_size(p: *const TfaParams, nb_block: c_int) -> c_int;
}
