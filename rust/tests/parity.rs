//! Parity of the shim with the reference chain on the SAME tfhe-rs keys and ciphertexts: the test a maintainer with cargo and a
//! B200 runs to turn "ciphertext-level parity unpinned"  into a measurement.
//!
//!   cargo test --release -- --nocapture --test-threads 1
//!
//! Each case runs the reference's algorithm through tfhe-rs on the CPU (the code of rostin79s/TFHE-AES src/server/, vendored by
//! the maintainer as the `reference` module below: copy `src/server` and `src/tables` of the reference next to this file) and the
//! shim on the GPU, then compares (1) the decrypted bytes (bit-exact, and against the `aes` crate as client.rs:163-171 does) and
//! (2) the noise variance of the outputs (ratio within [0.5, 2]).
use aes::cipher::{generic_array::GenericArray, BlockEncrypt, KeyInit};
use aes::Aes128;
use tfhe::integer::ciphertext::BaseRadixCiphertext;
use tfhe::integer::wopbs::WopbsKey;
use tfhe::integer::{gen_keys_radix, IntegerCiphertext, PublicKey, RadixClientKey, ServerKey};
use tfhe::shortint::parameters::parameters_wopbs::WopbsParameters;
use tfhe::shortint::parameters::*;
use tfhe::shortint::Ciphertext;

use tfhe_aes_b200 as shim;

#[path = "reference/mod.rs"]
mod reference; // the reference's src/server + src/tables, unmodified (see the module comment above)

type Byte = BaseRadixCiphertext<Ciphertext>;

/// client.rs:31-57
pub const PARAM_OPT: WopbsParameters = WopbsParameters {
    lwe_dimension: LweDimension(669),
    glwe_dimension: GlweDimension(4),
    polynomial_size: PolynomialSize(512),
    lwe_noise_distribution: DynamicDistribution::new_gaussian_from_std_dev(StandardDev(3.0517578125e-05)),
    glwe_noise_distribution: DynamicDistribution::new_gaussian_from_std_dev(StandardDev(3.162026630747649e-16)),
    pbs_base_log: DecompositionBaseLog(8),
    pbs_level: DecompositionLevelCount(5),
    ks_level: DecompositionLevelCount(6),
    ks_base_log: DecompositionBaseLog(2),
    pfks_level: DecompositionLevelCount(3),
    pfks_base_log: DecompositionBaseLog(12),
    pfks_noise_distribution: DynamicDistribution::new_gaussian_from_std_dev(StandardDev(3.162026630747649e-16)),
    cbs_level: DecompositionLevelCount(1),
    cbs_base_log: DecompositionBaseLog(15),
    message_modulus: MessageModulus(2),
    carry_modulus: CarryModulus(1),
    ciphertext_modulus: CiphertextModulus::new_native(),
    encryption_key_choice: EncryptionKeyChoice::Big,
};

struct Keys {
    cks: RadixClientKey,
    sks: ServerKey,
    wopbs: WopbsKey,
    pk: PublicKey,
}

fn keygen() -> Keys {
    // client.rs:106-107, :141
    let (cks, sks) = gen_keys_radix(PARAM_OPT, 8);
    let wopbs = WopbsKey::new_wopbs_key_only_for_wopbs(&cks, &sks);
    let pk = PublicKey::new(&cks);
    Keys { cks, sks, wopbs, pk }
}

fn encrypt_block(k: &Keys, block: [u8; 16]) -> Vec<Byte> {
    block.iter().map(|&b| k.cks.encrypt_without_padding(b as u64)).collect() // client.rs:126-138 (byte 0 = most significant)
}
fn decrypt_block(k: &Keys, ct: &[Byte]) -> [u8; 16] {
    let mut out = [0u8; 16];
    for (o, c) in out.iter_mut().zip(ct) {
        *o = k.cks.decrypt_without_padding::<u64>(c) as u8; // client.rs:154
    }
    out
}

/// signed phase error of every block of a byte: b - <a, s> - bit * 2^63 with the big LWE secret key
fn errors(k: &Keys, ct: &[Byte]) -> Vec<f64> {
    use tfhe::core_crypto::prelude::*;
    let (shortint_cks, _) = k.cks.as_ref().clone().into_raw_parts();
    let (glwe_sk, _lwe_sk, _params) = shortint_cks.into_raw_parts();
    let big_sk = glwe_sk.into_lwe_secret_key();
    let mut out = Vec::new();
    for byte in ct {
        for block in byte.blocks() {
            let pt = decrypt_lwe_ciphertext(&big_sk, &block.ct).0;
            let bit = pt.wrapping_add(1 << 62) >> 63;
            out.push(pt.wrapping_sub(bit << 63) as i64 as f64);
        }
    }
    out
}
fn variance(v: &[f64]) -> f64 {
    let m = v.iter().sum::<f64>() / v.len() as f64;
    v.iter().map(|x| (x - m) * (x - m)).sum::<f64>() / v.len() as f64
}

#[test]
fn many_sbox_matches_the_reference() {
    let k = keygen();
    let short = k.wopbs.clone().into_raw_parts();
    let (mut e_ref, mut e_gpu) = (Vec::new(), Vec::new());
    for x in (0u64..256).step_by(5) {
        for inv in [false, true] {
            let mut a = k.cks.encrypt_without_padding(x);
            let mut b = a.clone();
            let r = reference::sbox::sbox::many_sbox(&short, &mut a, inv);
            let g = shim::sbox::many_sbox(&short, &mut b, inv);
            assert_eq!(r.len(), g.len());
            for (cr, cg) in r.iter().zip(&g) {
                assert_eq!(k.cks.decrypt_without_padding::<u64>(cr), k.cks.decrypt_without_padding::<u64>(cg), "x = {x}, inv = {inv}");
            }
            e_ref.extend(errors(&k, &r));
            e_gpu.extend(errors(&k, &g));
        }
    }
    assert!(e_gpu.len() >= 1000);
    let ratio = variance(&e_gpu) / variance(&e_ref);
    println!("many_sbox noise: std ref {:.3e}, gpu {:.3e}, variance ratio {:.3}", variance(&e_ref).sqrt(), variance(&e_gpu).sqrt(), ratio);
    assert!(ratio > 0.5 && ratio < 2.0);
}

#[test]
fn gen_lut_is_bit_identical() {
    for nb in [8usize, 9] {
        let f = |x: u64| (x * 7 + 3) & ((1 << nb) - 1);
        let r = reference::sbox::gen_lut::gen_lut(2, 1, 512, nb, f);
        let g = shim::sbox::gen_lut(2, 1, 512, nb, f);
        assert_eq!(r.as_ref().lut().as_ref(), g.as_ref().lut().as_ref());
    }
}

#[test]
fn key_expansion_encrypt_decrypt_match_the_reference_and_the_aes_crate() {
    // main.rs:76-118 (the reference's commented-out test()): SP 800-38A F.1.1
    let k = keygen();
    let key: [u8; 16] = [0x2b, 0x7e, 0x15, 0x16, 0x28, 0xae, 0xd2, 0xa6, 0xab, 0xf7, 0x15, 0x88, 0x09, 0xcf, 0x4f, 0x3c];
    let pt: [u8; 16] = [0x6b, 0xc1, 0xbe, 0xe2, 0x2e, 0x40, 0x9f, 0x96, 0xe9, 0x3d, 0x7e, 0x11, 0x73, 0x93, 0x17, 0x2a];
    let srv_ref = reference::server::Server::new(k.pk.clone(), k.sks.clone(), k.wopbs.clone());
    let srv_gpu = shim::Server::new(k.pk.clone(), k.sks.clone(), k.wopbs.clone());
    let enc_key = encrypt_block(&k, key);
    let rk_ref = srv_ref.aes_key_expansion(&enc_key);
    let rk_gpu = srv_gpu.aes_key_expansion(&enc_key);
    for (a, b) in rk_ref.iter().zip(&rk_gpu) {
        assert_eq!(decrypt_block(&k, a), decrypt_block(&k, b));
    }
    let mut s_ref = encrypt_block(&k, pt);
    let mut s_gpu = s_ref.clone();
    srv_ref.aes_encrypt(&rk_ref, &mut s_ref);
    srv_gpu.aes_encrypt(&rk_gpu, &mut s_gpu);
    let mut want = GenericArray::clone_from_slice(&pt);
    Aes128::new(&GenericArray::from(key)).encrypt_block(&mut want);
    assert_eq!(decrypt_block(&k, &s_ref)[..], want[..]);
    assert_eq!(decrypt_block(&k, &s_gpu)[..], want[..]);
    let ratio = variance(&errors(&k, &s_gpu)) / variance(&errors(&k, &s_ref));
    println!("aes_encrypt output noise: variance ratio gpu / reference {ratio:.3} (128 samples)");
    srv_ref.aes_decrypt(&rk_ref, &mut s_ref);
    srv_gpu.aes_decrypt(&rk_gpu, &mut s_gpu);
    assert_eq!(decrypt_block(&k, &s_ref), pt);
    assert_eq!(decrypt_block(&k, &s_gpu), pt);
}

#[test]
fn ctr_counters_including_the_carry_fix() {
    // counters 0, 1, 255 agree with the reference; 256 and 1023 agree with the `aes` crate (the reference's add_scalar is
    // wrong for i >= 256, server.rs:181-182, so its own assert at client.rs:171 fails there)
    let k = keygen();
    let srv_ref = reference::server::Server::new(k.pk.clone(), k.sks.clone(), k.wopbs.clone());
    let srv_gpu = shim::Server::new(k.pk.clone(), k.sks.clone(), k.wopbs.clone());
    let rk = srv_gpu.aes_key_expansion(&encrypt_block(&k, [0u8; 16]));
    let iv = encrypt_block(&k, [0u8; 16]);
    for i in [0u128, 1, 255] {
        let mut a = iv.clone();
        let mut b = iv.clone();
        srv_ref.add_scalar(&mut a, i);
        srv_gpu.add_scalar(&mut b, i);
        assert_eq!(decrypt_block(&k, &a), decrypt_block(&k, &b), "counter {i}");
    }
    let cipher = Aes128::new(&GenericArray::from([0u8; 16]));
    for (first, n) in [(254u128, 3usize), (1022, 2)] {
        let out = srv_gpu.aes_ctr(&rk, &iv, first, n);
        for (j, block) in out.iter().enumerate() {
            let mut want = GenericArray::from((first + j as u128).to_be_bytes());
            cipher.encrypt_block(&mut want);
            assert_eq!(decrypt_block(&k, block)[..], want[..], "counter {}", first + j as u128);
        }
    }
}

#[test]
fn per_block_calls_from_rayon_workers() {
    // the reference's call pattern (main.rs:55-64): one block per worker on a shared &Server
    use rayon::prelude::*;
    let k = keygen();
    let srv = shim::Server::new(k.pk.clone(), k.sks.clone(), k.wopbs.clone());
    let rk = srv.aes_key_expansion(&encrypt_block(&k, [7u8; 16]));
    let iv = encrypt_block(&k, [0u8; 16]);
    let outs: Vec<Vec<Byte>> = (0..16u128).into_par_iter().map(|i| {
        let mut state = iv.clone();
        srv.add_scalar(&mut state, i);
        srv.aes_encrypt(&rk, &mut state);
        state
    }).collect();
    let cipher = Aes128::new(&GenericArray::from([7u8; 16]));
    for (i, block) in outs.iter().enumerate() {
        let mut want = GenericArray::from((i as u128).to_be_bytes());
        cipher.encrypt_block(&mut want);
        assert_eq!(decrypt_block(&k, block)[..], want[..]);
    }
}
