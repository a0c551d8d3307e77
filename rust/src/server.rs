//! `Server` of the reference (src/server/server.rs:24-282) on the B200 engine: identical public signatures.
use std::collections::HashMap;
use std::ffi::{c_int, c_void, CStr};
use std::sync::{Arc, Mutex, OnceLock};

use tfhe::integer::ciphertext::BaseRadixCiphertext;
use tfhe::integer::wopbs::WopbsKey;
use tfhe::integer::{IntegerCiphertext, PublicKey, ServerKey};
use tfhe::shortint::Ciphertext;

use crate::ffi::{self, TfaParams};
use crate::flatten::{export_keys, flatten_state, params_of, unflatten};

/// One `tfa_ctx`: a GPU context with the prepared keys.  Thread-safe: the library serialises (and, for the per-block entry
/// points, coalesces) concurrent calls, so `&Engine` can be shared by rayon workers as `&Server` is in main.rs:55-64.
pub struct Engine {
    pub(crate) ctx: *mut c_void,
    pub(crate) params: TfaParams,
}
unsafe impl Send for Engine {}
unsafe impl Sync for Engine {}

impl Engine {
    pub(crate) fn check(&self, rc: c_int) {
        if rc != 0 {
            // the reference panics on every error (unwrap / assert!); so does the shim
            let msg = unsafe { CStr::from_ptr(ffi::tfa_last_error(self.ctx)) }.to_string_lossy().into_owned();
            panic!("tfhe_aes_b200 error {rc}: {msg}");
        }
    }
    /// `device` = CUDA device index (one process per GPU in the multi-GPU setting)
    pub fn new(wopbs_key_short: &tfhe::shortint::wopbs::WopbsKey, device: i32) -> Arc<Engine> {
        let params = params_of(wopbs_key_short);
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { ffi::tfa_ctx_create(&params, device, std::ptr::null_mut(), &mut ctx) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(ffi::tfa_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            panic!("tfhe_aes_b200: cannot create a context ({rc}): {msg}");
        }
        let engine = Engine { ctx, params };
        let keys = export_keys(wopbs_key_short);
        engine.check(unsafe { ffi::tfa_ctx_load_keys(ctx, keys.bsk.as_ptr(), keys.ksk.as_ptr(), keys.pfpksk.as_ptr()) });
        Arc::new(engine)
    }
}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::tfa_ctx_destroy(self.ctx) }
    }
}

/// Engines by key identity, so that the free functions of the `sbox` module, which only receive `&WopbsKey` (sbox.rs:46,68,
/// many_wopbs.rs:31), find the context `Server::new` created for that key (or create one on first use).
fn registry() -> &'static Mutex<HashMap<usize, Arc<Engine>>> {
    static R: OnceLock<Mutex<HashMap<usize, Arc<Engine>>>> = OnceLock::new();
    R.get_or_init(|| Mutex::new(HashMap::new()))
}
fn key_id(k: &tfhe::shortint::wopbs::WopbsKey) -> usize {
    k.cbs_pfpksk.as_ref().as_ptr() as usize // the PFPKSK buffer lives as long as the key and is unique to it
}
pub(crate) fn engine_for(k: &tfhe::shortint::wopbs::WopbsKey) -> Arc<Engine> {
    let mut r = registry().lock().unwrap();
    r.entry(key_id(k))
        .or_insert_with(|| {
            let device = std::env::var("TFHE_AES_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            Engine::new(k, device)
        })
        .clone()
}

pub struct Server {
    public_key: PublicKey,
    #[allow(dead_code)]
    sks: ServerKey,
    #[allow(dead_code)]
    wopbs_key: WopbsKey,
    wopbs_key_short: tfhe::shortint::wopbs::WopbsKey,
    engine: Arc<Engine>,
}

type Byte = BaseRadixCiphertext<Ciphertext>;

impl Server {
    /// server.rs:32-35.  The keys go to the GPU here, once: BSK back to the standard domain, KSK / PFPKSK in ABI level order.
    pub fn new(public_key: PublicKey, sks: ServerKey, wopbs_key: WopbsKey) -> Self {
        let wopbs_key_short = wopbs_key.clone().into_raw_parts();
        let engine = engine_for(&wopbs_key_short);
        Server { public_key, sks, wopbs_key, wopbs_key_short, engine }
    }

    fn flatten_round_keys(rks: &[Vec<Byte>]) -> Vec<u64> {
        let mut out = Vec::new();
        for rk in rks {
            out.extend(flatten_state(rk));
        }
        out
    }

    /// server.rs:39-64
    pub fn aes_encrypt(&self, encrypted_round_keys: &Vec<Vec<Byte>>, state: &mut Vec<Byte>) {
        let rk = Self::flatten_round_keys(encrypted_round_keys);
        let mut st = flatten_state(state);
        self.engine.check(unsafe { ffi::tfa_aes_encrypt(self.engine.ctx, rk.as_ptr(), st.as_mut_ptr(), 1) });
        *state = unflatten(&st, 8, &state[0]);
    }
    /// README.md:58 spelling
    pub fn aes_encryption(&self, encrypted_round_keys: &Vec<Vec<Byte>>, state: &mut Vec<Byte>) {
        self.aes_encrypt(encrypted_round_keys, state)
    }

    /// server.rs:67-105
    pub fn aes_decrypt(&self, encrypted_round_keys: &Vec<Vec<Byte>>, state: &mut Vec<Byte>) {
        let rk = Self::flatten_round_keys(encrypted_round_keys);
        let mut st = flatten_state(state);
        self.engine.check(unsafe { ffi::tfa_aes_decrypt(self.engine.ctx, rk.as_ptr(), st.as_mut_ptr(), 1) });
        *state = unflatten(&st, 8, &state[0]);
    }
    /// README.md:59 spelling
    pub fn aes_decryption(&self, encrypted_round_keys: &Vec<Vec<Byte>>, state: &mut Vec<Byte>) {
        self.aes_decrypt(encrypted_round_keys, state)
    }

    /// server.rs:107-167.  RCON is encrypted with the public key exactly as server.rs:139-140 does and handed to the engine.
    pub fn aes_key_expansion(&self, key: &Vec<Byte>) -> Vec<Vec<Byte>> {
        use crate::sbox::RCON;
        let flat_key = flatten_state(key);
        let rcon: Vec<Byte> = RCON.iter().map(|&r| self.public_key.encrypt_radix_without_padding(r as u64, 8)).collect();
        let flat_rcon = flatten_state(&rcon);
        let mut rk = vec![0u64; 11 * flat_key.len()];
        self.engine.check(unsafe { ffi::tfa_aes_key_expansion(self.engine.ctx, flat_key.as_ptr(), flat_rcon.as_ptr(), rk.as_mut_ptr()) });
        rk.chunks_exact(flat_key.len()).map(|r| unflatten(r, 8, &key[0])).collect()
    }

    /// server.rs:172-275 (the engine uses `i & 0xFF` in the low-byte LUTs: server.rs:181-182 is wrong for i >= 256)
    pub fn add_scalar(&self, state: &mut Vec<Byte>, i: u128) {
        let mut st = flatten_state(state);
        let ctr = [i as u64, (i >> 64) as u64];
        self.engine.check(unsafe { ffi::tfa_add_scalar(self.engine.ctx, st.as_mut_ptr(), ctr.as_ptr(), 1) });
        *state = unflatten(&st, 8, &state[0]);
    }

    /// main.rs:55-64 in one call: out[b] = AES(iv + first + b).  A GPU is filled by batches, not by one block per worker.
    pub fn aes_ctr(&self, encrypted_round_keys: &Vec<Vec<Byte>>, encrypted_iv: &Vec<Byte>, first: u128, number_of_outputs: usize) -> Vec<Vec<Byte>> {
        let rk = Self::flatten_round_keys(encrypted_round_keys);
        let iv = flatten_state(encrypted_iv);
        let mut out = vec![0u64; number_of_outputs * iv.len()];
        self.engine.check(unsafe {
            ffi::tfa_aes_ctr(self.engine.ctx, rk.as_ptr(), iv.as_ptr(), first as u64, (first >> 64) as u64, number_of_outputs as c_int, out.as_mut_ptr())
        });
        out.chunks_exact(iv.len()).map(|s| unflatten(s, 8, &encrypted_iv[0])).collect()
    }

    /// the shortint key, for the `sbox` module functions (server.rs passes `&self.wopbs_key_short`)
    pub fn wopbs_key_short(&self) -> &tfhe::shortint::wopbs::WopbsKey {
        &self.wopbs_key_short
    }
}
