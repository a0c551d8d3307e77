//! tfhe-rs containers <-> the flat little-endian `u64` arrays of the C ABI (include/tfhe_aes_b200.h, "layouts").
//!
//! Ciphertexts are zero-copy in spirit (every tfhe-rs container on this path is a flat `Vec<u64>`); keys are exported once,
//! in `Server::new`:
//!   * the bootstrap key is kept by tfhe-rs in the Fourier domain in a plan-dependent coefficient order that must not be
//!     consumed, so it is brought back to the standard domain polynomial by polynomial with tfhe-rs' own inverse transform;
//!   * the keyswitch key and the private functional packing keyswitch keys are standard-domain already; their
//!     decomposition levels are re-ordered into the order the ABI defines (level 1 = scale q/beta first), whatever order
//!     this tfhe-rs version stores them in (`LEVELS_STORED_HIGHEST_FIRST`).
use tfhe::core_crypto::fft_impl::fft64::crypto::bootstrap::FourierLweBootstrapKeyView;
use tfhe::core_crypto::fft_impl::fft64::math::fft::Fft;
use tfhe::core_crypto::prelude::*;
use tfhe::integer::ciphertext::BaseRadixCiphertext;
use tfhe::integer::IntegerCiphertext;
use tfhe::shortint::parameters::{Degree, NoiseLevel};
use tfhe::shortint::server_key::ShortintBootstrappingKey;
use tfhe::shortint::wopbs::WopbsKey;
use tfhe::shortint::Ciphertext;

use crate::ffi::TfaParams;

/// tfhe-rs 0.10+ stores the level blocks of keyswitch-type keys and the level matrices of GGSW ciphertexts from the highest
/// decomposition level (smallest scale q/beta^l) to level 1, so that they zip with the decomposition iterator, which yields
/// level l first.  The ABI of libtfhe_aes_b200 defines level 1 first.  `tests/parity.rs` fails on the first S-box if this
/// constant does not match the tfhe-rs version in Cargo.toml.
pub const LEVELS_STORED_HIGHEST_FIRST: bool = true;

/// `WopbsParameters` (client.rs:31-57) -> `tfa_params`
pub fn params_of(key: &WopbsKey) -> TfaParams {
    let p = &key.param;
    let std_of = |d: DynamicDistribution<u64>| d.gaussian_std_dev().0;
    TfaParams {
        lwe_dim: p.lwe_dimension.0 as u32,
        glwe_dim: p.glwe_dimension.0 as u32,
        poly_size: p.polynomial_size.0 as u32,
        pbs_base_log: p.pbs_base_log.0 as u32,
        pbs_level: p.pbs_level.0 as u32,
        ks_base_log: p.ks_base_log.0 as u32,
        ks_level: p.ks_level.0 as u32,
        pfks_base_log: p.pfks_base_log.0 as u32,
        pfks_level: p.pfks_level.0 as u32,
        cbs_base_log: p.cbs_base_log.0 as u32,
        cbs_level: p.cbs_level.0 as u32,
        message_modulus: p.message_modulus.0 as u32,
        carry_modulus: p.carry_modulus.0 as u32,
        _pad: 0,
        lwe_std: std_of(p.lwe_noise_distribution),
        glwe_std: std_of(p.glwe_noise_distribution),
        pfks_std: std_of(p.pfks_noise_distribution),
    }
}

/// one radix ciphertext (a byte: 8 blocks, block j = bit j, LSB first) -> `[nblocks][lw]`
pub fn flatten_radix(ct: &BaseRadixCiphertext<Ciphertext>, out: &mut Vec<u64>) {
    for block in ct.blocks() {
        out.extend_from_slice(block.ct.as_ref());
    }
}

/// a state / a round key (16 bytes) -> `[16][8][lw]`
pub fn flatten_state(state: &[BaseRadixCiphertext<Ciphertext>]) -> Vec<u64> {
    let mut out = Vec::with_capacity(state.len() * 8 * 2049);
    for byte in state {
        flatten_radix(byte, &mut out);
    }
    out
}

/// Re-wraps `[nbytes][nblocks][lw]` as radix ciphertexts exactly as many_wopbs.rs:87-114 does: every LWE becomes a shortint
/// block with `Degree = message_modulus - 1`, `NoiseLevel::NOMINAL` and the moduli / PBS order of the matching input block.
pub fn unflatten(flat: &[u64], nblocks: usize, template: &BaseRadixCiphertext<Ciphertext>) -> Vec<BaseRadixCiphertext<Ciphertext>> {
    let lw = template.blocks()[0].ct.lwe_size().0;
    let modulus = template.blocks()[0].ct.ciphertext_modulus();
    assert_eq!(flat.len() % (nblocks * lw), 0, "flat ciphertext buffer is not a whole number of radix ciphertexts");
    flat.chunks_exact(nblocks * lw)
        .map(|radix| {
            let blocks: Vec<Ciphertext> = radix
                .chunks_exact(lw)
                .enumerate()
                .map(|(j, lwe)| {
                    // a 9-block counter-add radix borrows the attributes of block 0 for its carry block (server.rs:216-222)
                    let like = &template.blocks()[j.min(template.blocks().len() - 1)];
                    Ciphertext::new(
                        LweCiphertextOwned::from_container(lwe.to_vec(), modulus),
                        Degree::new(like.message_modulus.0 - 1),
                        NoiseLevel::NOMINAL,
                        like.message_modulus,
                        like.carry_modulus,
                        like.pbs_order,
                    )
                })
                .collect();
            BaseRadixCiphertext::from_blocks(blocks)
        })
        .collect()
}

/// Standard-domain bootstrap key `[n][level 1..l][row][(k+1) * N]` from the Fourier key of the server key: every Fourier
/// polynomial goes through tfhe-rs' own `add_backward_as_torus` into a zeroed standard polynomial, so the plan-dependent
/// coefficient order never leaves tfhe-rs.  The round trip costs an error of about 2^-53 relative to 2^63 per coefficient,
/// far below the key's own noise (sigma = 3.16e-16 * 2^64 = 2^12.5).
pub fn export_bsk_standard(bsk: FourierLweBootstrapKeyView<'_>) -> Vec<u64> {
    let n = bsk.input_lwe_dimension().0;
    let glwe_size = bsk.glwe_size().0;
    let poly = bsk.polynomial_size();
    let levels = bsk.decomposition_level_count().0;
    let fft = Fft::new(poly);
    let fft = fft.as_view();
    let mut mem = tfhe::core_crypto::commons::computation_buffers::ComputationBuffers::new();
    mem.resize(fft.backward_scratch().unwrap().unaligned_bytes_required());
    let ggsw_words = levels * glwe_size * glwe_size * poly.0;
    let mut out = vec![0u64; n * ggsw_words];
    for (i, ggsw) in bsk.into_ggsw_iter().enumerate() {
        for matrix in ggsw.into_levels() {
            let level = matrix.decomposition_level().0; // 1..=levels, whatever the storage order
            for (row_index, row) in matrix.into_rows().enumerate() {
                let fourier_polys = row.data(); // glwe_size Fourier polynomials of N/2 complex each
                let half = poly.to_fourier_polynomial_size().0;
                for col in 0..glwe_size {
                    let off = i * ggsw_words + (((level - 1) * glwe_size + row_index) * glwe_size + col) * poly.0;
                    let standard = PolynomialMutView::from_container(&mut out[off..off + poly.0]);
                    let fourier = tfhe::core_crypto::fft_impl::fft64::math::polynomial::FourierPolynomialView {
                        data: &fourier_polys[col * half..(col + 1) * half],
                    };
                    fft.add_backward_as_torus(standard, fourier, mem.stack());
                }
            }
        }
    }
    out
}

/// Level blocks of one keyswitch-type key element `[levels][width]`, copied into ABI order (level 1 first).
fn copy_levels_abi_order(src: &[u64], levels: usize, width: usize, dst: &mut [u64]) {
    for slot in 0..levels {
        let level = if LEVELS_STORED_HIGHEST_FIRST { levels - slot } else { slot + 1 };
        dst[(level - 1) * width..level * width].copy_from_slice(&src[slot * width..(slot + 1) * width]);
    }
}

/// Keyswitch key big -> small: `[k*N][level 1..l][n+1]`
pub fn export_ksk(ksk: &LweKeyswitchKeyOwned<u64>) -> Vec<u64> {
    let levels = ksk.decomposition_level_count().0;
    let width = ksk.output_lwe_size().0;
    let src = ksk.as_ref();
    let mut out = vec![0u64; src.len()];
    for (s, d) in src.chunks_exact(levels * width).zip(out.chunks_exact_mut(levels * width)) {
        copy_levels_abi_order(s, levels, width, d);
    }
    out
}

/// Circuit-bootstrap PFPKSK list: `[k+1 keys][k*N+1 input elements][level 1..l][(k+1)*N]`
pub fn export_pfpksk(list: &LwePrivateFunctionalPackingKeyswitchKeyListOwned<u64>) -> Vec<u64> {
    let levels = list.decomposition_level_count().0;
    let width = list.output_glwe_size().0 * list.output_polynomial_size().0;
    let src = list.as_ref();
    let mut out = vec![0u64; src.len()];
    for (s, d) in src.chunks_exact(levels * width).zip(out.chunks_exact_mut(levels * width)) {
        copy_levels_abi_order(s, levels, width, d);
    }
    out
}

/// All three keys of the shortint `WopbsKey` in ABI layout.  `new_wopbs_key_only_for_wopbs` (client.rs:107) makes
/// `wopbs_server_key` and `pbs_server_key` clones of one server key, so one BSK and one KSK serve the whole chain.
pub struct ExportedKeys {
    pub bsk: Vec<u64>,
    pub ksk: Vec<u64>,
    pub pfpksk: Vec<u64>,
}

pub fn export_keys(key: &WopbsKey) -> ExportedKeys {
    let sks = &key.wopbs_server_key;
    let bsk = match &sks.bootstrapping_key {
        ShortintBootstrappingKey::Classic(bsk) => export_bsk_standard(bsk.as_view()),
        // the reference silently returns zero ciphertexts for this variant (many_wopbs.rs:83-84); refuse instead
        ShortintBootstrappingKey::MultiBit { .. } => panic!("tfhe_aes_b200: multi-bit bootstrap keys are not supported"),
    };
    ExportedKeys { bsk, ksk: export_ksk(&sks.key_switching_key), pfpksk: export_pfpksk(&key.cbs_pfpksk) }
}
