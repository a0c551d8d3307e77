"""Clear AES-128/192/256 written from FIPS-197 with a hard-coded S-box: the expected values of the parity tests, of
bench.py and of smoke().  Independent of the product (tfhe-aes_b200/) and of the oracle (oracle/): neither is imported
here.  When the `cryptography` package is importable, `aes_encrypt_block` is cross-checked against it once per process."""

SBOX = bytes.fromhex(
    "637c777bf26b6fc53001672bfed7ab76ca82c97dfa5947f0add4a2af9ca472c0b7fd9326363ff7cc34a5e5f171d8311504c723c31896059a071280e2eb27b275"
    "09832c1a1b6e5aa0523bd6b329e32f8453d100ed20fcb15b6acbbe394a4c58cfd0efaafb434d338545f9027f503c9fa851a3408f929d38f5bcb6da2110fff3d2"
    "cd0c13ec5f974417c4a77e3d645d197360814fdc222a908846eeb814de5e0bdbe0323a0a4906245cc2d3ac629195e479e7c8376d8dd54ea96c56f4ea657aae08"
    "ba78252e1ca6b4c6e8dd741f4bbd8b8a703eb5664803f60e613557b986c11d9ee1f8981169d98e949b1e87e9ce5528df8ca1890dbfe6426841992d0fb054bb16")
INV_SBOX = bytes(SBOX.index(i) for i in range(256))
RCON = (0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36)


def xtime(a):
    return ((a << 1) ^ 0x1B) & 0xFF if a & 0x80 else a << 1


def gmul(a, m):
    r = 0
    while m:
        if m & 1:
            r ^= a
        a = xtime(a)
        m >>= 1
    return r


def shift_rows(s, inverse=False):
    """state index = 4*column + row (FIPS-197 §3.4); row r is rotated left by r columns (right when inverse)."""
    sign = -1 if inverse else 1
    return bytes(s[4 * ((c + sign * r) % 4) + r] for c in range(4) for r in range(4))


def mix_columns(s, inverse=False):
    m = (14, 11, 13, 9) if inverse else (2, 3, 1, 1)
    out = bytearray(16)
    for c in range(4):
        col = s[4 * c:4 * c + 4]
        for r in range(4):
            out[4 * c + r] = gmul(col[r], m[0]) ^ gmul(col[(r + 1) % 4], m[1]) ^ gmul(col[(r + 2) % 4], m[2]) ^ gmul(col[(r + 3) % 4], m[3])
    return bytes(out)


def xor(a, b):
    return bytes(x ^ y for x, y in zip(a, b))


def round_keys(key):
    """FIPS-197 §5.2 key expansion for 16/24/32-byte keys: list of Nr+1 round keys of 16 bytes."""
    nk = len(key) // 4
    nr = nk + 6
    w = [key[4 * i:4 * i + 4] for i in range(nk)]
    for i in range(nk, 4 * (nr + 1)):
        t = w[i - 1]
        if i % nk == 0:
            t = bytes(SBOX[b] for b in t[1:] + t[:1])
            t = bytes([t[0] ^ RCON[i // nk - 1]]) + t[1:]
        elif nk > 6 and i % nk == 4:
            t = bytes(SBOX[b] for b in t)
        w.append(xor(w[i - nk], t))
    return [b"".join(w[4 * r:4 * r + 4]) for r in range(nr + 1)]


def aes_round(state, rk):
    """one middle round: SubBytes, ShiftRows, MixColumns, AddRoundKey (server.rs:44-55)"""
    return xor(mix_columns(shift_rows(bytes(SBOX[b] for b in state))), rk)


def aes_encrypt_block(key, block):
    rks = round_keys(key)
    s = xor(block, rks[0])
    for rk in rks[1:-1]:
        s = aes_round(s, rk)
    return xor(shift_rows(bytes(SBOX[b] for b in s)), rks[-1])


def aes_decrypt_block(key, block):
    rks = round_keys(key)
    s = xor(block, rks[-1])
    for rk in reversed(rks[1:-1]):
        s = mix_columns(xor(bytes(INV_SBOX[b] for b in shift_rows(s, True)), rk), True)
    return xor(bytes(INV_SBOX[b] for b in shift_rows(s, True)), rks[0])


def ctr_block(key, iv, i):
    """AES(key, iv + i mod 2^128), the reference's CTR keystream block i (main.rs:55-64, client.rs:147-175)"""
    return aes_encrypt_block(key, ((iv + i) % (1 << 128)).to_bytes(16, "big"))


def _self_check():
    assert aes_encrypt_block(bytes(16), bytes(16)).hex() == "66e94bd4ef8a2c3b884cfa59ca342b2e"
    assert aes_encrypt_block(bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c"), bytes.fromhex("6bc1bee22e409f96e93d7e117393172a")).hex() == "3ad77bb40d7a3660a89ecaf32466ef97"
    # FIPS-197 Appendix C.2 / C.3
    pt = bytes.fromhex("00112233445566778899aabbccddeeff")
    assert aes_encrypt_block(bytes(range(24)), pt).hex() == "dda97ca4864cdfe06eaf70a0ec0d7191"
    assert aes_encrypt_block(bytes(range(32)), pt).hex() == "8ea2b7ca516745bfeafc49904b496089"
    assert aes_decrypt_block(bytes(range(32)), bytes.fromhex("8ea2b7ca516745bfeafc49904b496089")) == pt
    try:
        from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    except Exception:
        return
    import os
    for n in (16, 24, 32):
        k, b = os.urandom(n), os.urandom(16)
        assert Cipher(algorithms.AES(k), modes.ECB()).encryptor().update(b) == aes_encrypt_block(k, b)


_self_check()
