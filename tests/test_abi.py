"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/tfhe_aes_b200.h, the host-only entry points work, and compute entry points fail loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tfhe_aes_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tfhe_aes_b200.h but not exported"
    for n in ("tfa_aes_key_expansion", "tfa_aes_encrypt", "tfa_aes_decrypt", "tfa_aes_encryption", "tfa_aes_decryption",
              "tfa_add_scalar", "tfa_many_wopbs", "tfa_sbox", "tfa_many_sbox", "tfa_gen_lut"):
        assert n in names


def test_param_opt_matches_reference(pkg):
    lib = pkg.load_library()
    p = pkg.Params()
    lib.tfa_param_opt(C.byref(p))
    ref = pkg.param_opt()
    for f, _ in pkg.Params._fields_:
        assert getattr(p, f) == getattr(ref, f)
    # client.rs:31-57
    assert (p.lwe_dim, p.glwe_dim, p.poly_size) == (669, 4, 512)
    assert (p.pbs_base_log, p.pbs_level, p.ks_base_log, p.ks_level) == (8, 5, 2, 6)
    assert (p.pfks_base_log, p.pfks_level, p.cbs_base_log, p.cbs_level) == (12, 3, 15, 1)
    assert (p.message_modulus, p.carry_modulus) == (2, 1)


def test_gen_lut_host_entry_point(pkg):
    p = pkg.param_opt()
    assert pkg.lut_size(p, 8) == 512 and pkg.lut_size(p, 9) == 512 and pkg.lut_size(p, 10) == 1024
    lut = pkg.gen_lut(p, 8, lambda x: pkg.SBOX[x])
    assert lut.shape == (8, 512)
    for j in range(8):
        assert np.array_equal(lut[j, :256] >> np.uint64(63), np.array([(pkg.SBOX[x] >> j) & 1 for x in range(256)], dtype=np.uint64))
    assert np.all((lut & np.uint64((1 << 63) - 1)) == 0)


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the context cannot be created: the product never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.TfaError) as ei:
        pkg.Engine(pkg.param_opt())
    assert ei.value.code == 2  # TFA_ERR_CUDA


def test_unsupported_parameters_are_rejected(pkg):
    p = pkg.param_opt()
    p.poly_size = 1024
    with pytest.raises(pkg.TfaError) as ei:
        pkg.Engine(p)
    assert ei.value.code == 1  # TFA_ERR_PARAM


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tfhe-aes_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) and "emu" not in f:
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "tfhe_oracle" not in txt and "import orc" not in txt, f


def test_client_generator_is_chacha20(pkg):
    """The client harness draws masks, key bits and noise from the ChaCha20 key stream (csrc/client.cu): RFC 7539 §2.3.2
    block-function vector (key 00..1f, counter word 1, nonce 00:00:00:09 00:00:00:4a 00:00:00:00)."""
    lib = pkg.load_library()
    key = (C.c_uint32 * 8)(*[int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)])
    out = (C.c_uint64 * 8)()
    lib.tfa_rng_block.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.tfa_rng_block.restype = None
    # RFC words 12..15 = 00000001 09000000 4a000000 00000000 -> 64-bit counter = words 12,13 and 64-bit nonce = words 14,15
    lib.tfa_rng_block(key, 0x000000004a000000, (0x09000000 << 32) | 1, out)
    stream = b"".join(int(v).to_bytes(8, "little") for v in out)
    assert stream.hex() == ("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
                            "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")


def test_rust_shim_binds_only_declared_symbols(pkg):
    """rust/src/ffi.rs (the shim a maintainer compiles against tfhe-rs, which cannot be built in this image) declares exactly
    entry points that the header declares and the library exports, with matching argument counts."""
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "tfhe_aes_b200.h")).read(), flags=re.S)
    lib = pkg.load_library()
    decls = re.findall(r"pub fn (tfa_[a-z0-9_]+)\((.*?)\)", ffi, flags=re.S)
    assert len(decls) >= 12
    for name, args in decls:
        assert hasattr(lib, name), name
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, hdr, flags=re.S)
        assert m, f"{name} is not declared in the header"
        n_rust = len([a for a in args.split(",") if a.strip()])
        n_c = len([a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"])
        assert n_rust == n_c, (name, n_rust, n_c)
    # the reference's surface is all there (server.rs:32,39,67,107,172; sbox.rs:46,68; many_wopbs.rs:31; gen_lut.rs:9)
    srv = open(os.path.join(ROOT, "rust", "src", "server.rs")).read()
    for sig in ("pub fn new(public_key: PublicKey, sks: ServerKey, wopbs_key: WopbsKey) -> Self", "pub fn aes_encrypt(&self,", "pub fn aes_decrypt(&self,",
                "pub fn aes_key_expansion(&self,", "pub fn add_scalar(&self,"):
        assert sig in srv, sig
    sb = open(os.path.join(ROOT, "rust", "src", "sbox.rs")).read()
    for sig in ("pub fn gen_lut<F>(message_mod: usize, carry_mod: usize, poly_size: usize, nb_block: usize, f: F) -> IntegerWopbsLUT",
                "pub fn many_wopbs_without_padding(ct_in: &mut BaseRadixCiphertext<Ciphertext>, wopbs_key_short: &WopbsKey, luts: Vec<IntegerWopbsLUT>)",
                "pub fn sbox(", "pub fn many_sbox(wopbs_key_short: &WopbsKey, ct_in: &mut BaseRadixCiphertext<Ciphertext>, inv: bool)"):
        assert sig in sb, sig
    for f in ("Cargo.toml", "build.rs", "README.md", "src/lib.rs", "src/flatten.rs", "tests/parity.rs"):
        assert os.path.exists(os.path.join(ROOT, "rust", f)), f
    assert "unimplemented!" not in srv + sb + open(os.path.join(ROOT, "rust", "src", "flatten.rs")).read()
