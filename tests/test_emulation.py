"""CPU emulation of the CTA-level CUDA code (tfhe-aes_b200/csrc/emu.cu runs the very functions the
kernels run, thread by thread) against the oracle: FFT, inverse FFT and whole CMux steps."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tfhe-aes_b200", "libtfhe_aes_emu.so")


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(EMU):
        pytest.skip("emulation library not built (run __graft_entry__.build())")
    return C.CDLL(EMU)


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_forward_fft_matches_oracle(emu, oracle_test):
    rng = np.random.default_rng(0)
    for _ in range(4):
        poly = rng.integers(0, 2 ** 64, 512, dtype=np.uint64)
        out = np.zeros((256, 2))
        emu.emu_fft_forward_torus(P(poly), P(out))
        ref = oracle_test.fft_forward_torus(poly)
        assert np.abs(out[:, 0] + 1j * out[:, 1] - ref).max() <= 2e-15 * np.abs(ref).max()


def test_fft_roundtrip(emu):
    rng = np.random.default_rng(1)
    poly = rng.integers(0, 2 ** 64, 512, dtype=np.uint64)
    rt = np.zeros(512, dtype=np.uint64)
    emu.emu_fft_roundtrip(P(poly), P(rt))
    with np.errstate(over="ignore"):
        d = (rt - poly).astype(np.int64)
    assert np.abs(d).max() < 2 ** 14  # 64-bit input through a 53-bit mantissa: ~2^11 rounding + FFT error


def _rotate_diff(acc, rot, N=512):
    j = np.arange(N)
    s = (j - rot) % (2 * N)
    with np.errstate(over="ignore"):
        x = acc[:, s % N]
        x = np.where(s >= N, np.uint64(0) - x, x)
        return x - acc


@pytest.mark.parametrize("K,G,bl,lv", [(1, 1, 8, 5), (1, 4, 8, 5), (1, 2, 15, 1)])
def test_cmux_step_matches_oracle(emu, orc, oracle_test, K, G, bl, lv):
    o = oracle_test
    rng = np.random.default_rng(K * 100 + G)
    if (bl, lv) == (8, 5):
        ggsw = o.bsk().reshape(o.n, lv, K + 1, (K + 1) * 512)[3].copy()
    else:
        ggsw = o.circuit_bootstrap_boolean(o.encrypt_lwe_small(np.array([1 << 63], dtype=np.uint64))[0])
    acc = rng.integers(0, 2 ** 64, (G, K + 1, 512), dtype=np.uint64)
    rot = rng.integers(0, 1024, G).astype(np.int32)
    exp = np.stack([o.external_product_add(ggsw, bl, lv, _rotate_diff(acc[g], int(rot[g])).ravel(), acc[g].ravel()).reshape(K + 1, 512)
                    for g in range(G)])
    got = acc.copy()
    assert emu.emu_cmux_rotate(K, G, bl, lv, P(np.ascontiguousarray(ggsw)), P(rot), P(got)) == 0
    with np.errstate(over="ignore"):
        d = np.abs((got - exp).astype(np.int64)).max()
    # FFT rounding budget per CMux: ~2^25 for the PBS product, ~2^31 for the level-1 product
    assert d < (2 ** 29 if lv == 5 else 2 ** 35)


def test_cmux_step_param_opt_shape(emu, orc, oracle_opt):
    """K = 4, G = 3: the production instantiation of the PBS kernel."""
    o = oracle_opt
    rng = np.random.default_rng(5)
    ggsw = o.bsk().reshape(o.n, 5, 5, 5 * 512)[11].copy()
    acc = rng.integers(0, 2 ** 64, (3, 5, 512), dtype=np.uint64)
    rot = np.array([1, 513, 1000], dtype=np.int32)
    exp = np.stack([o.external_product_add(ggsw, 8, 5, _rotate_diff(acc[g], int(rot[g])).ravel(), acc[g].ravel()).reshape(5, 512) for g in range(3)])
    got = acc.copy()
    assert emu.emu_cmux_rotate(4, 3, 8, 5, P(np.ascontiguousarray(ggsw)), P(rot), P(got)) == 0
    with np.errstate(over="ignore"):
        assert np.abs((got - exp).astype(np.int64)).max() < 2 ** 30
