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


def _balanced_digits_85(x):
    """The (2^8, 5) recoding of cmux_core.cuh (decomp85_*): digits in [-128, 127], least significant
    (level 5) first, with sum_j d_j 2^(8j+24) = closest_representable(x) mod 2^64."""
    with np.errstate(over="ignore"):
        y = (x + np.uint64(0x8080808080800000)) ^ np.uint64(0x8080808080000000)
    return [((y >> np.uint64(24 + 8 * j)) & np.uint64(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64) for j in range(5)]


def test_balanced_digits_represent_the_closest_value(orc):
    """Same value as the reference's decomposer (oracle restatement), different tie rule."""
    rng = np.random.default_rng(3)
    x = rng.integers(0, 2 ** 64, 4096, dtype=np.uint64)
    x[:4] = [0, (1 << 64) - 1, 0x0000008000800000, 0x7F7F7F7F7F800000]
    mine = _balanced_digits_85(x)
    with np.errstate(over="ignore"):
        val = sum(d.astype(np.uint64) << np.uint64(24 + 8 * j) for j, d in enumerate(mine))
        for i in range(x.size):
            ref = orc.decompose(int(x[i]), 8, 5)      # index 0 = level 5 (least significant)
            assert int(val[i]) == sum(int(d) << (24 + 8 * j) for j, d in enumerate(ref)) % 2 ** 64
            assert all(-128 <= int(m[i]) <= 127 for m in mine)


def _negacyclic_matrix(g):
    """T with (T @ d)[i] = sum_j d[j] g[i-j] (negacyclic), exact modulo 2^64 in uint64 arithmetic."""
    N = g.size
    i, j = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    with np.errstate(over="ignore"):
        return np.where(i >= j, g[(i - j) % N], np.uint64(0) - g[(i - j) % N])


def _exact_external_product_add_85(ggsw, diff, acc, K):
    """acc + GGSW (x) diff with the kernel's digit recoding, exact integer arithmetic modulo 2^64."""
    out = acc.copy()
    digs = [_balanced_digits_85(diff[r]) for r in range(K + 1)]   # [row][j] -> 512 digits, j = 0 is level 5
    with np.errstate(over="ignore"):
        for lev in range(1, 6):
            for r in range(K + 1):
                d = digs[r][5 - lev].astype(np.uint64)
                for c in range(K + 1):
                    out[c] += _negacyclic_matrix(ggsw[lev - 1, r, c * 512:(c + 1) * 512]) @ d
    return out


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
    if lv == 5:
        # (2^8, 5): the kernel recodes exact ties differently from the reference's decomposer, which swaps the
        # masks for other, equally valid ones.  (a) raw words against exact integer arithmetic with the kernel's
        # digits: only FFT rounding remains (~2^25 per CMux); (b) against the oracle on the phase.
        for g in range(G):
            exact = _exact_external_product_add_85(ggsw.reshape(5, K + 1, (K + 1) * 512), _rotate_diff(acc[g], int(rot[g])), acc[g], K)
            with np.errstate(over="ignore"):
                assert np.abs((got[g] - exact).astype(np.int64)).max() < 2 ** 29
                dp = (o.glwe_phase(got[g].ravel()) - o.glwe_phase(exp[g].ravel())).astype(np.int64)
            assert np.abs(dp).max() < 2 ** 32
        return
    with np.errstate(over="ignore"):
        d = np.abs((got - exp).astype(np.int64)).max()
    assert d < 2 ** 35  # FFT rounding budget of the level-1 product (~2^31)


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
    for g in range(3):
        exact = _exact_external_product_add_85(ggsw, _rotate_diff(acc[g], int(rot[g])), acc[g], 4)
        with np.errstate(over="ignore"):
            assert np.abs((got[g] - exact).astype(np.int64)).max() < 2 ** 30          # FFT rounding only
            dp = (o.glwe_phase(got[g].ravel()) - o.glwe_phase(exp[g].ravel())).astype(np.int64)
        assert np.abs(dp).max() < 2 ** 32                                              # same message, noise-level difference
