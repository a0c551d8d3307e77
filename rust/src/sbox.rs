//! The reference's `sbox` module (src/server/sbox/{gen_lut,sbox,many_wopbs}.rs) and constants, same signatures.
use tfhe::integer::ciphertext::BaseRadixCiphertext;
use tfhe::integer::wopbs::{CiphertextCount, IntegerWopbsLUT, PlaintextCount};
use tfhe::integer::IntegerCiphertext;
use tfhe::shortint::wopbs::WopbsKey;
use tfhe::shortint::Ciphertext;

use crate::ffi;
use crate::flatten::{flatten_radix, unflatten};
use crate::server::engine_for;

/// key_expansion_utils.rs:10-12
pub const RCON: [u8; 10] = [0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36];

/// sbox.rs:20-42 (GF(2^8), polynomial 0x11B), in the clear
pub fn mul2(x: u8) -> u8 {
    (x << 1) ^ if x & 0x80 != 0 { 0x1B } else { 0 }
}
pub fn mul3(x: u8) -> u8 {
    mul2(x) ^ x
}
pub fn mul9(x: u8) -> u8 {
    mul2(mul2(mul2(x))) ^ x
}
pub fn mul11(x: u8) -> u8 {
    mul2(mul2(mul2(x)) ^ x) ^ x
}
pub fn mul13(x: u8) -> u8 {
    mul2(mul2(mul2(x) ^ x)) ^ x
}
pub fn mul14(x: u8) -> u8 {
    mul2(mul2(mul2(x) ^ x) ^ x)
}

/// gen_lut.rs:9-42 through `tfa_gen_lut` (host only): `lut[j][index] = bit_j(f(index mod 2^nb_block)) << 63`, with
/// `max(2^(nb_block * log_basis), poly_size)` entries per row.
pub fn gen_lut<F>(message_mod: usize, carry_mod: usize, poly_size: usize, nb_block: usize, f: F) -> IntegerWopbsLUT
where
    F: Fn(u64) -> u64,
{
    let p = ffi::TfaParams {
        lwe_dim: 0, glwe_dim: 0, poly_size: poly_size as u32, pbs_base_log: 0, pbs_level: 0, ks_base_log: 0, ks_level: 0,
        pfks_base_log: 0, pfks_level: 0, cbs_base_log: 0, cbs_level: 0,
        message_modulus: message_mod as u32, carry_modulus: carry_mod as u32, _pad: 0, lwe_std: 0.0, glwe_std: 0.0, pfks_std: 0.0,
    };
    let lut_size = unsafe { ffi::tfa_lut_size(&p, nb_block as i32) };
    assert!(lut_size > 0, "tfhe_aes_b200: unsupported LUT shape");
    let lut_size = lut_size as usize;
    let log_basis = ((message_mod * carry_mod) as f64).log2() as usize;
    let table: Vec<u64> = (0..1u64 << (nb_block * log_basis)).map(&f).collect();
    let mut flat = vec![0u64; nb_block * lut_size];
    let rc = unsafe { ffi::tfa_gen_lut(&p, nb_block as i32, table.as_ptr(), flat.as_mut_ptr()) };
    assert_eq!(rc, 0, "tfhe_aes_b200: tfa_gen_lut failed");
    let mut lut = IntegerWopbsLUT::new(PlaintextCount(lut_size), CiphertextCount(nb_block));
    for block in 0..nb_block {
        lut[block].copy_from_slice(&flat[block * lut_size..(block + 1) * lut_size]);
    }
    lut
}

/// many_wopbs.rs:31-116: one circuit bootstrap of `ct_in`'s blocks, one vertical packing per LUT.
pub fn many_wopbs_without_padding(ct_in: &mut BaseRadixCiphertext<Ciphertext>, wopbs_key_short: &WopbsKey, luts: Vec<IntegerWopbsLUT>) -> Vec<BaseRadixCiphertext<Ciphertext>> {
    let engine = engine_for(wopbs_key_short);
    let nblocks = ct_in.blocks().len();
    let mut flat_in = Vec::new();
    flatten_radix(ct_in, &mut flat_in);
    let lw = flat_in.len() / nblocks;
    let mut flat_luts = Vec::new();
    for lut in &luts {
        assert_eq!(lut.as_ref().output_ciphertext_count().0, nblocks, "LUT rows must match the blocks of the input");
        flat_luts.extend_from_slice(lut.as_ref().lut().as_ref());
    }
    let mut out = vec![0u64; luts.len() * nblocks * lw];
    engine.check(unsafe { ffi::tfa_many_wopbs(engine.ctx, flat_in.as_ptr(), 1, nblocks as i32, flat_luts.as_ptr(), luts.len() as i32, out.as_mut_ptr()) });
    unflatten(&out, nblocks, ct_in)
}

/// sbox.rs:46-63: in-place S-box (or inverse S-box) of one encrypted byte
pub fn sbox(_wopbs_key: &tfhe::integer::wopbs::WopbsKey, wopbs_key_short: &WopbsKey, ct_in: &mut BaseRadixCiphertext<Ciphertext>, inv: bool) {
    let engine = engine_for(wopbs_key_short);
    let mut flat = Vec::new();
    flatten_radix(ct_in, &mut flat);
    engine.check(unsafe { ffi::tfa_sbox(engine.ctx, flat.as_mut_ptr(), 1, inv as i32) });
    *ct_in = unflatten(&flat, 8, ct_in).pop().unwrap();
}

/// sbox.rs:68-97: {S, 2S, 3S}(x) for encryption, {9, 11, 13, 14}·x for decryption, one circuit bootstrap
pub fn many_sbox(wopbs_key_short: &WopbsKey, ct_in: &mut BaseRadixCiphertext<Ciphertext>, inv: bool) -> Vec<BaseRadixCiphertext<Ciphertext>> {
    let engine = engine_for(wopbs_key_short);
    let mut flat = Vec::new();
    flatten_radix(ct_in, &mut flat);
    let nluts = if inv { 4 } else { 3 };
    let mut out = vec![0u64; nluts * flat.len()];
    engine.check(unsafe { ffi::tfa_many_sbox(engine.ctx, flat.as_ptr(), 1, inv as i32, out.as_mut_ptr()) });
    unflatten(&out, 8, ct_in)
}
