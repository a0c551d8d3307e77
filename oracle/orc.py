"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product path (tfhe-aes_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libtfhe_oracle.so")


class Params(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "lwe_dim", "glwe_dim", "poly_size", "pbs_base_log", "pbs_level", "ks_base_log", "ks_level",
        "pfks_base_log", "pfks_level", "cbs_base_log", "cbs_level", "message_modulus", "carry_modulus",
        "_pad")] + [("lwe_std", C.c_double), ("glwe_std", C.c_double), ("pfks_std", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "_pad"}


def param_opt():
    """PARAM_OPT, /root/reference/src/client/client.rs:31-57."""
    return Params(669, 4, 512, 8, 5, 2, 6, 12, 3, 15, 1, 2, 1, 0,
                  3.0517578125e-05, 3.162026630747649e-16, 3.162026630747649e-16)


def param_test():
    """Small functional parameter set (same decomposition bases as PARAM_OPT, n=24, k=1, tiny noise).

    The reference's authors suggest "lower security parameters to see correctness faster"
    (main.rs:75) but provide none; this is ours.  Not secure — tests only."""
    return Params(24, 1, 512, 8, 5, 2, 6, 12, 3, 15, 1, 2, 1, 0, 2.0 ** -24, 2.0 ** -50, 2.0 ** -50)


def param_test2():
    """Like param_test but 2-bit message blocks (message_modulus 4): exercises the PBS loop of the
    general extract_bits and LUTs spanning several polynomials (CMux tree)."""
    return Params(24, 1, 512, 8, 5, 2, 6, 12, 3, 15, 1, 4, 1, 0, 2.0 ** -24, 2.0 ** -50, 2.0 ** -50)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("tfhe_oracle.cpp", "tfhe_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_create.restype = C.c_void_p
        for name in ("orc_lwe_sk", "orc_glwe_sk", "orc_bsk", "orc_ksk", "orc_pfpksk", "orc_sbox_table"):
            getattr(_lib, name).restype = C.c_void_p
        _lib.orc_gf_mul.restype = C.c_uint8
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def u64(*shape):
    return np.zeros(shape, dtype=np.uint64)


class Oracle:
    def __init__(self, params, seed=None):
        self.L = lib()
        self.params = params
        self.h = C.c_void_p(self.L.orc_create(C.byref(params)))
        self.n, self.k, self.N = params.lwe_dim, params.glwe_dim, params.poly_size
        self.big = self.k * self.N
        self.lw = self.big + 1
        self.gsz = (self.k + 1) * self.N
        if seed is not None:
            self.keygen(seed)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # -- keys ------------------------------------------------------------------------------------
    def keygen(self, seed):
        self.L.orc_keygen(self.h, C.c_uint64(seed))

    def _view(self, fn, n):
        ptr = getattr(self.L, fn)(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n,))

    def lwe_sk(self):
        return self._view("orc_lwe_sk", self.n)

    def glwe_sk(self):
        return self._view("orc_glwe_sk", self.big)

    def bsk(self):
        p = self.params
        return self._view("orc_bsk", self.n * p.pbs_level * (self.k + 1) * self.gsz)

    def ksk(self):
        p = self.params
        return self._view("orc_ksk", self.big * p.ks_level * (self.n + 1))

    def pfpksk(self):
        p = self.params
        return self._view("orc_pfpksk", (self.k + 1) * (self.big + 1) * p.pfks_level * self.gsz)

    # -- client ----------------------------------------------------------------------------------
    def seed_encryption(self, seed):
        self.L.orc_seed_encryption(self.h, C.c_uint64(seed))

    def encrypt_bits(self, bits):
        bits = np.ascontiguousarray(bits, dtype=np.uint8).ravel()
        out = u64(len(bits), self.lw)
        self.L.orc_encrypt_bits(self.h, _p(bits), len(bits), _p(out))
        return out

    def encrypt_bytes(self, data):
        data = np.frombuffer(bytes(data), dtype=np.uint8).copy()
        out = u64(len(data), 8, self.lw)
        self.L.orc_encrypt_bytes(self.h, _p(data), len(data), _p(out))
        return out

    def trivial_bytes(self, data):
        data = np.frombuffer(bytes(data), dtype=np.uint8).copy()
        out = u64(len(data), 8, self.lw)
        self.L.orc_trivial_bytes(self.h, _p(data), len(data), _p(out))
        return out

    def decrypt_bits(self, ct, with_err=False):
        ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1, self.lw)
        bits = np.zeros(len(ct), dtype=np.uint8)
        err = np.zeros(len(ct), dtype=np.int64)
        self.L.orc_decrypt_bits(self.h, _p(ct), len(ct), _p(bits), _p(err))
        return (bits, err) if with_err else bits

    def decrypt_bytes(self, ct):
        ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1, 8, self.lw)
        out = np.zeros(len(ct), dtype=np.uint8)
        self.L.orc_decrypt_bytes(self.h, _p(ct), len(ct), _p(out))
        return bytes(out)

    def encrypt_lwe_small(self, plain):
        plain = np.ascontiguousarray(plain, dtype=np.uint64).ravel()
        out = u64(len(plain), self.n + 1)
        self.L.orc_encrypt_lwe_small(self.h, _p(plain), len(plain), _p(out))
        return out

    def phase_small(self, ct):
        ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1, self.n + 1)
        out = u64(len(ct))
        self.L.orc_phase_small(self.h, _p(ct), len(ct), _p(out))
        return out

    def phase_big(self, ct):
        ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1, self.lw)
        out = u64(len(ct))
        self.L.orc_phase_big(self.h, _p(ct), len(ct), _p(out))
        return out

    def glwe_phase(self, glwe):
        glwe = np.ascontiguousarray(glwe, dtype=np.uint64).reshape(-1, self.gsz)
        out = u64(len(glwe), self.N)
        self.L.orc_glwe_phase(self.h, _p(glwe), len(glwe), _p(out))
        return out

    # -- primitives ------------------------------------------------------------------------------
    def fft_forward_torus(self, poly):
        poly = np.ascontiguousarray(poly, dtype=np.uint64)
        out = np.zeros((self.N // 2, 2), dtype=np.float64)
        self.L.orc_fft_forward_torus(self.h, _p(poly), _p(out))
        return out[:, 0] + 1j * out[:, 1]

    def fft_forward_integer(self, poly):
        poly = np.ascontiguousarray(poly, dtype=np.int64)
        out = np.zeros((self.N // 2, 2), dtype=np.float64)
        self.L.orc_fft_forward_integer(self.h, _p(poly), _p(out))
        return out[:, 0] + 1j * out[:, 1]

    def fft_add_backward_torus(self, fourier, poly):
        f = np.ascontiguousarray(np.stack([fourier.real, fourier.imag], axis=1), dtype=np.float64)
        poly = np.array(poly, dtype=np.uint64)
        self.L.orc_fft_add_backward_torus(self.h, _p(f), _p(poly))
        return poly

    def keyswitch(self, ct):
        ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1, self.lw)
        out = u64(len(ct), self.n + 1)
        self.L.orc_keyswitch(self.h, _p(ct), len(ct), _p(out))
        return out

    def bootstrap(self, ct, lut):
        ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(-1, self.n + 1)
        lut = np.ascontiguousarray(lut, dtype=np.uint64)
        assert lut.shape == (self.N,)
        out = u64(len(ct), self.lw)
        self.L.orc_bootstrap(self.h, _p(ct), len(ct), _p(lut), _p(out))
        return out

    def extract_bits(self, ct, delta_log, nbits):
        ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(self.lw)
        out = u64(nbits, self.n + 1)
        self.L.orc_extract_bits(self.h, _p(ct), delta_log, nbits, _p(out))
        return out

    def pfks(self, key_index, lwe):
        lwe = np.ascontiguousarray(lwe, dtype=np.uint64).reshape(self.lw)
        out = u64(self.gsz)
        self.L.orc_pfks(self.h, key_index, _p(lwe), _p(out))
        return out

    def circuit_bootstrap_boolean(self, lwe, delta_log=63):
        lwe = np.ascontiguousarray(lwe, dtype=np.uint64).reshape(self.n + 1)
        out = u64(self.params.cbs_level, self.k + 1, self.gsz)
        self.L.orc_circuit_bootstrap_boolean(self.h, _p(lwe), delta_log, _p(out))
        return out

    def vertical_packing(self, lut_polys, ggsw_std):
        lut_polys = np.ascontiguousarray(lut_polys, dtype=np.uint64).reshape(-1, self.N)
        ggsw_std = np.ascontiguousarray(ggsw_std, dtype=np.uint64)
        nggsw = ggsw_std.shape[0]
        out = u64(self.lw)
        self.L.orc_vertical_packing(self.h, _p(lut_polys), len(lut_polys), _p(ggsw_std), nggsw, _p(out))
        return out

    def external_product_add(self, ggsw_std, base_log, level, glwe_in, glwe_acc):
        ggsw_std = np.ascontiguousarray(ggsw_std, dtype=np.uint64)
        glwe_in = np.ascontiguousarray(glwe_in, dtype=np.uint64)
        acc = np.array(glwe_acc, dtype=np.uint64)
        self.L.orc_external_product_add(self.h, _p(ggsw_std), base_log, level, _p(glwe_in), _p(acc))
        return acc

    # -- sbox module -----------------------------------------------------------------------------
    def lut_size(self, nb_block):
        return self.L.orc_lut_size(C.byref(self.params), nb_block)

    def gen_lut(self, nb_block, table):
        table = np.ascontiguousarray(table, dtype=np.uint64)
        out = u64(nb_block, self.lut_size(nb_block))
        self.L.orc_gen_lut(C.byref(self.params), nb_block, _p(table), _p(out))
        return out

    def many_wopbs(self, ct_in, luts):
        ct_in = np.ascontiguousarray(ct_in, dtype=np.uint64).reshape(-1, self.lw)
        nb = len(ct_in)
        luts = np.ascontiguousarray(luts, dtype=np.uint64)
        L = luts.shape[0]
        assert luts.shape[1] == nb
        out = u64(L, nb, self.lw)
        self.L.orc_many_wopbs(self.h, _p(ct_in), nb, _p(luts), L, _p(out))
        return out

    def sbox(self, byte_ct, inv=False):
        b = np.array(byte_ct, dtype=np.uint64).reshape(8, self.lw)
        self.L.orc_sbox(self.h, _p(b), int(inv))
        return b

    def many_sbox(self, byte_ct, inv=False):
        b = np.ascontiguousarray(byte_ct, dtype=np.uint64).reshape(8, self.lw)
        out = u64(4 if inv else 3, 8, self.lw)
        self.L.orc_many_sbox(self.h, _p(b), int(inv), _p(out))
        return out

    # -- Server ----------------------------------------------------------------------------------
    def aes_key_expansion(self, key_ct, rcon_ct=None):
        key_ct = np.ascontiguousarray(key_ct, dtype=np.uint64).reshape(16, 8, self.lw)
        out = u64(11, 16, 8, self.lw)
        rp = _p(np.ascontiguousarray(rcon_ct, dtype=np.uint64)) if rcon_ct is not None else None
        self.L.orc_aes_key_expansion(self.h, _p(key_ct), rp, _p(out))
        return out

    def aes_encrypt(self, rk, state):
        rk = np.ascontiguousarray(rk, dtype=np.uint64)
        st = np.array(state, dtype=np.uint64).reshape(16, 8, self.lw)
        self.L.orc_aes_encrypt(self.h, _p(rk), _p(st))
        return st

    def aes_decrypt(self, rk, state):
        rk = np.ascontiguousarray(rk, dtype=np.uint64)
        st = np.array(state, dtype=np.uint64).reshape(16, 8, self.lw)
        self.L.orc_aes_decrypt(self.h, _p(rk), _p(st))
        return st

    def add_scalar(self, state, counter, faithful=False):
        st = np.array(state, dtype=np.uint64).reshape(16, 8, self.lw)
        self.L.orc_add_scalar(self.h, _p(st), C.c_uint64(counter & (2 ** 64 - 1)), C.c_uint64(counter >> 64), int(faithful))
        return st

    def add_round_key(self, state, rk):
        st = np.array(state, dtype=np.uint64).reshape(16, 8, self.lw)
        rk = np.ascontiguousarray(rk, dtype=np.uint64)
        self.L.orc_add_round_key(self.h, _p(st), _p(rk))
        return st

    def mix_columns(self, mul_state):
        m = np.ascontiguousarray(mul_state, dtype=np.uint64).reshape(16, 3, 8, self.lw)
        out = u64(16, 8, self.lw)
        self.L.orc_mix_columns(self.h, _p(m), _p(out))
        return out

    def inv_mix_columns(self, mul_state):
        m = np.ascontiguousarray(mul_state, dtype=np.uint64).reshape(16, 4, 8, self.lw)
        out = u64(16, 8, self.lw)
        self.L.orc_inv_mix_columns(self.h, _p(m), _p(out))
        return out

    def shift_rows(self, state, inverse=False):
        st = np.array(state, dtype=np.uint64).reshape(16, 8, self.lw)
        self.L.orc_shift_rows(self.h, _p(st), int(inverse))
        return st


# -- clear AES (FIPS-197) ---------------------------------------------------------------------------
def clear_aes_encrypt(key: bytes, block: bytes) -> bytes:
    out = (C.c_uint8 * 16)()
    lib().orc_clear_aes_encrypt(bytes(key), bytes(block), out)
    return bytes(out)


def clear_aes_decrypt(key: bytes, block: bytes) -> bytes:
    out = (C.c_uint8 * 16)()
    lib().orc_clear_aes_decrypt(bytes(key), bytes(block), out)
    return bytes(out)


def clear_round_keys(key: bytes) -> bytes:
    out = (C.c_uint8 * 176)()
    lib().orc_clear_aes_key_expansion(bytes(key), out)
    return bytes(out)


def sbox_table(inv=False):
    ptr = lib().orc_sbox_table(int(inv))
    return bytes(C.cast(ptr, C.POINTER(C.c_uint8 * 256)).contents)


def gf_mul(x, m):
    return lib().orc_gf_mul(C.c_uint8(x), m)


def decompose(x, base_log, level):
    out = np.zeros(level, dtype=np.int64)
    lib().orc_decompose(C.c_uint64(int(x)), base_log, level, _p(out))
    return out
