"""Generates tests/golden/golden.json from the CPU oracle (seeded, PARAM_TEST) and from FIPS-197.

The reference holds no ciphertext-level fixture; what it does hold — the four
SP 800-38A F.1.1 blocks of main.rs:78-95 and the always-on `aes`-crate comparison — is recorded here
as plaintext-level known answers, next to digests of oracle ciphertext outputs that pin the oracle
itself against accidental change.

Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import orc  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    g = {}
    g["sp800_38a_f11"] = {
        "key": "2b7e151628aed2a6abf7158809cf4f3c",
        "plain": ["6bc1bee22e409f96e93d7e117393172a", "ae2d8a571e03ac9c9eb76fac45af8e51",
                  "30c81c46a35ce411e5fbc1191a0a52ef", "f69f2445df4f9b17ad2b417be66c3710"],
        "cipher": ["3ad77bb40d7a3660a89ecaf32466ef97", "f5d3d58503b9699de785895a96fdbaaf",
                   "43b1cd7f598ece23881b00e3ed030688", "7b0c785e27e8ad3f8223207104725dd4"],
    }
    g["ctr_key0_iv0"] = {"0": "66e94bd4ef8a2c3b884cfa59ca342b2e", "1": "58e2fccefa7e3061367f1d57a4e7455a",
                         "255": "f70ddef93ba62588242a0e67d0d645e0", "256": "fb56cc09b680b1d07c5a52149e29f07c",
                         "1023": "09836c44648b1f262e7b8bb6310e3c73"}
    g["sbox_first_row"] = orc.sbox_table()[:16].hex()
    g["inv_sbox_first_row"] = orc.sbox_table(True)[:16].hex()
    # decomposition vectors: x, base_log, level -> digits (level `level` first)
    rng = np.random.default_rng(2024)
    dec = []
    for bl, lv in ((8, 5), (2, 6), (12, 3), (15, 1)):
        for x in list(rng.integers(0, 2 ** 64, 6, dtype=np.uint64)) + [np.uint64(0), np.uint64(2 ** 64 - 1), np.uint64(1 << 63), np.uint64((1 << (64 - bl * lv)) >> 1)]:
            dec.append({"x": int(x), "base_log": bl, "level": lv, "digits": [int(d) for d in orc.decompose(int(x), bl, lv)]})
    g["decompose"] = dec
    # seeded oracle outputs on PARAM_TEST (digests) + a small explicit vector
    o = orc.Oracle(orc.param_test(), seed=1)
    g["param_test_seed1"] = {
        "lwe_sk": "".join(str(int(b)) for b in o.lwe_sk()),
        "bsk_sha256": digest(o.bsk()), "ksk_sha256": digest(o.ksk()), "pfpksk_sha256": digest(o.pfpksk()),
    }
    o.seed_encryption(42)
    ct = o.encrypt_bytes(bytes([0x53]))
    ks = o.keyswitch(ct[0])
    g["param_test_seed1"]["encrypt_0x53_sha256"] = digest(ct)
    g["param_test_seed1"]["keyswitch_sha256"] = digest(ks)
    g["param_test_seed1"]["keyswitch_first_lwe"] = [int(v) for v in ks[0]]
    g["param_test_seed1"]["pfks0_sha256"] = digest(o.pfks(0, ct[0, 0]))
    lut = o.gen_lut(8, np.frombuffer(orc.sbox_table(), dtype=np.uint8).astype(np.uint64))
    g["gen_lut_sbox_sha256"] = digest(lut)
    g["gen_lut_sbox_row0_head"] = [int(v >> np.uint64(63)) for v in lut[0, :32]]
    lut9 = o.gen_lut(9, np.array([((x & 0xFF) + (x >> 8) + 0x7F) % 256 for x in range(512)], dtype=np.uint64))
    g["gen_lut_add9_sha256"] = digest(lut9)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
