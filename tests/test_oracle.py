"""CPU tests of the oracle against every known answer the reference holds for this path
(main.rs:78-95 SP 800-38A vectors, the `aes`-crate check client.rs:171) and the committed fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "golden.json")) as f:
    G = json.load(f)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_clear_aes_known_answers(orc):
    v = G["sp800_38a_f11"]
    key = bytes.fromhex(v["key"])
    for p, c in zip(v["plain"], v["cipher"]):
        assert orc.clear_aes_encrypt(key, bytes.fromhex(p)).hex() == c
        assert orc.clear_aes_decrypt(key, bytes.fromhex(c)).hex() == p
    for ctr, c in G["ctr_key0_iv0"].items():
        assert orc.clear_aes_encrypt(bytes(16), int(ctr).to_bytes(16, "big")).hex() == c
    try:  # independent implementation, if present in the image
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
        rng = np.random.default_rng(0)
        for _ in range(10):  # main.rs:120-141: ten random (key, plaintext) pairs
            k, p = bytes(rng.integers(0, 256, 16, dtype=np.uint8)), bytes(rng.integers(0, 256, 16, dtype=np.uint8))
            enc = Cipher(algorithms.AES(k), modes.ECB()).encryptor()
            assert orc.clear_aes_encrypt(k, p) == enc.update(p) + enc.finalize()
    except ImportError:
        pass


def test_tables_and_gf_mul(orc, pkg):
    assert orc.sbox_table()[:16].hex() == G["sbox_first_row"] == "637c777bf26b6fc53001672bfed7ab76"
    assert orc.sbox_table(True)[:16].hex() == G["inv_sbox_first_row"] == "52096ad53036a538bf40a39e81f3d7fb"
    assert orc.sbox_table() == pkg.SBOX and orc.sbox_table(True) == pkg.INV_SBOX
    for x in range(256):
        for m, f in ((2, pkg.mul2), (3, pkg.mul3), (9, pkg.mul9), (11, pkg.mul11), (13, pkg.mul13), (14, pkg.mul14)):
            assert orc.gf_mul(x, m) == f(x)
    assert pkg.mul2(0x80) == 0x1B and pkg.mul3(0x53) == (pkg.mul2(0x53) ^ 0x53)  # sbox.rs:20-32


def test_decomposition_vectors_and_properties(orc):
    for v in G["decompose"]:
        d = orc.decompose(v["x"], v["base_log"], v["level"])
        assert [int(t) for t in d] == v["digits"]
    rng = np.random.default_rng(1)
    for bl, lv in ((8, 5), (2, 6), (12, 3), (15, 1)):
        r = 64 - bl * lv
        for x in rng.integers(0, 2 ** 64, 200, dtype=np.uint64):
            x = int(x)
            d = [int(t) for t in orc.decompose(x, bl, lv)]
            assert all(-(1 << (bl - 1)) <= t <= (1 << (bl - 1)) for t in d)
            # recomposition: digits come out level `lv` first
            rec = sum(t << (64 - bl * (lv - i)) for i, t in enumerate(d)) % 2 ** 64
            closest = (((x >> r) + ((x >> (r - 1)) & 1)) << r) % 2 ** 64
            assert rec == closest


def test_seeded_oracle_is_pinned(orc, oracle_test):
    g = G["param_test_seed1"]
    o = oracle_test
    assert "".join(str(int(b)) for b in o.lwe_sk()) == g["lwe_sk"]
    assert digest(o.bsk()) == g["bsk_sha256"] and digest(o.ksk()) == g["ksk_sha256"] and digest(o.pfpksk()) == g["pfpksk_sha256"]
    o.seed_encryption(42)
    ct = o.encrypt_bytes(bytes([0x53]))
    assert digest(ct) == g["encrypt_0x53_sha256"]
    ks = o.keyswitch(ct[0])
    assert digest(ks) == g["keyswitch_sha256"]
    assert [int(v) for v in ks[0]] == g["keyswitch_first_lwe"]
    assert digest(o.pfks(0, ct[0, 0])) == g["pfks0_sha256"]


def test_gen_lut_matches_reference_definition(orc, pkg, oracle_test):
    sb = np.frombuffer(orc.sbox_table(), dtype=np.uint8).astype(np.uint64)
    lut = oracle_test.gen_lut(8, sb)
    assert digest(lut) == G["gen_lut_sbox_sha256"]
    assert [int(v >> np.uint64(63)) for v in lut[0, :32]] == G["gen_lut_sbox_row0_head"]
    # gen_lut.rs:19-22: 256 entries padded to poly_size 512; entries >= 256 repeat (index mod 256)
    assert lut.shape == (8, 512) and np.array_equal(lut[:, :256], lut[:, 256:])
    for j in range(8):
        assert np.array_equal(lut[j, :256] >> np.uint64(63), (sb >> np.uint64(j)) & np.uint64(1))
    # the product's host gen_lut is the same function
    assert np.array_equal(pkg.gen_lut(pkg.param_test(), 8, lambda x: int(sb[x])), lut)
    t9 = np.array([((x & 0xFF) + (x >> 8) + 0x7F) % 256 for x in range(512)], dtype=np.uint64)
    assert digest(oracle_test.gen_lut(9, t9)) == G["gen_lut_add9_sha256"]
    assert np.array_equal(pkg.gen_lut(pkg.param_test(), 9, t9), oracle_test.gen_lut(9, t9))


def test_fft_negacyclic_product(oracle_test):
    """pointwise product of two spectra = negacyclic product."""
    o = oracle_test
    rng = np.random.default_rng(2)
    a = rng.integers(-128, 129, 512)
    b = rng.integers(-2 ** 30, 2 ** 30, 512)  # exact in f64: 2^7 * 2^30 * 512 < 2^53
    fa, fb = o.fft_forward_integer(a), o.fft_forward_integer(b)
    got = o.fft_add_backward_torus(fa * fb, np.zeros(512, dtype=np.uint64)).astype(np.int64)
    full = np.convolve(a.astype(object), b.astype(object))
    exp = [int(full[i]) - (int(full[i + 512]) if i + 512 < len(full) else 0) for i in range(512)]
    assert [int(v) for v in got] == exp


def test_sbox_chain_on_test_params(orc, oracle_test):
    o = oracle_test
    data = bytes([0x00, 0x53, 0xA7, 0xFF])
    ct = o.encrypt_bytes(data)
    sb = orc.sbox_table()
    for i, b in enumerate(data):
        assert o.decrypt_bytes(o.sbox(ct[i]))[0] == sb[b]
        assert o.decrypt_bytes(o.sbox(ct[i], inv=True))[0] == orc.sbox_table(True)[b]
        m = o.decrypt_bytes(o.many_sbox(ct[i]))
        assert m == bytes([sb[b], orc.gf_mul(sb[b], 2), orc.gf_mul(sb[b], 3)])
        m = o.decrypt_bytes(o.many_sbox(ct[i], inv=True))
        assert m == bytes(orc.gf_mul(b, k) for k in (9, 11, 13, 14))


def test_reference_test_function_on_test_params(orc, oracle_test):
    """main.rs:76-118 (test()): key expansion -> encrypt -> decrypt on the SP 800-38A vectors."""
    o = oracle_test
    v = G["sp800_38a_f11"]
    key = bytes.fromhex(v["key"])
    rk = o.aes_key_expansion(o.encrypt_bytes(key))
    assert b"".join(o.decrypt_bytes(rk[r]) for r in range(11)) == orc.clear_round_keys(key)
    for p, c in list(zip(v["plain"], v["cipher"]))[:2]:
        st = o.aes_encrypt(rk, o.encrypt_bytes(bytes.fromhex(p)))
        assert o.decrypt_bytes(st).hex() == c
        assert o.decrypt_bytes(o.aes_decrypt(rk, st)).hex() == p


def test_add_scalar_and_the_reference_defect(oracle_test):
    o = oracle_test
    iv = (2 ** 128 - 300).to_bytes(16, "big")
    for ctr in (0, 1, 255, 256, 1023):
        got = o.decrypt_bytes(o.add_scalar(o.encrypt_bytes(iv), ctr))
        assert got == ((int.from_bytes(iv, "big") + ctr) % 2 ** 128).to_bytes(16, "big")
    # server.rs:181-182 uses the whole counter in the low-byte LUT: iv=0, i=256 gives 0x200, not 0x100
    bad = o.decrypt_bytes(o.add_scalar(o.encrypt_bytes(bytes(16)), 256, faithful=True))
    assert int.from_bytes(bad, "big") == 0x200
    ok = o.decrypt_bytes(o.add_scalar(o.encrypt_bytes(bytes(16)), 255, faithful=True))
    assert int.from_bytes(ok, "big") == 255


def test_general_forms_on_two_bit_blocks(orc, oracle_test2):
    """extract_bits with its PBS loop (2 bits per block) and a LUT over 4 blocks."""
    o = oracle_test2
    sb = np.frombuffer(orc.sbox_table(), dtype=np.uint8).astype(np.uint64)
    lut = o.gen_lut(4, sb)
    for v in (0x00, 0xC9):
        ct = np.zeros((4, o.lw), dtype=np.uint64)
        for blk in range(4):
            ct[blk] = o.encrypt_bits([0])[0]
            ct[blk, -1] += np.uint64(((v >> (2 * blk)) & 3) << 62)
        bits = o.extract_bits(ct[1], 62, 2)
        ph = o.phase_small(bits)
        dec = ((ph + np.uint64(1 << 62)) >> np.uint64(63)) & np.uint64(1)
        assert (int(dec[0]) << 1 | int(dec[1])) == (v >> 2) & 3
        out = o.many_wopbs(ct, lut[None])
        ph = o.phase_big(out[0])
        got = sum(int(((p + np.uint64(1 << 61)) >> np.uint64(62)) & np.uint64(3)) << (2 * b) for b, p in enumerate(ph))
        assert got == int(sb[v])


def test_sbox_on_param_opt(orc, oracle_opt):
    """PARAM_OPT (client.rs:31-57): one many_sbox, noise far from the decision boundary."""
    o = oracle_opt
    ct = o.encrypt_bytes(bytes([0x53]))
    out = o.many_sbox(ct[0])
    assert o.decrypt_bytes(out) == bytes([0xED, orc.gf_mul(0xED, 2), orc.gf_mul(0xED, 3)])
    _, err = o.decrypt_bits(out, with_err=True)
    assert np.abs(err.astype(np.float64)).max() * np.sqrt(5) < 2.0 ** 60
