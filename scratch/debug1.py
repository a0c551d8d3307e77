import sys, os, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
from conftest import load_pkg, torus_absdiff
import orc
pkg = load_pkg()
o = orc.Oracle(orc.param_test(), seed=1)
e = pkg.Engine(pkg.param_test()); e.load_keys(o.bsk(), o.ksk(), o.pfpksk())
rng = np.random.default_rng(0)
lut = rng.integers(0, 2**64, 512, dtype=np.uint64)
def run(lwe, tag):
    got, ref = e.bootstrap(lwe, lut), o.bootstrap(lwe, lut)
    d = [torus_absdiff(got[i], ref[i]) for i in range(len(lwe))]
    print(tag, ["2^%.1f" % np.log2(max(x,1)) for x in d])
# a: zero mask
for cnt in (1, 4, 9):
    lwe = np.zeros((cnt, o.n+1), dtype=np.uint64); lwe[:, -1] = rng.integers(0, 2**64, cnt, dtype=np.uint64)
    run(lwe, f"zero-mask cnt={cnt}")
# b: single nonzero mask element
for pos in (0, 5, 23):
    lwe = np.zeros((1, o.n+1), dtype=np.uint64); lwe[0, pos] = np.uint64(37 << 54); lwe[0,-1] = np.uint64(3<<54)
    run(lwe, f"single mask pos={pos}")
lwe = np.zeros((4, o.n+1), dtype=np.uint64); lwe[:, 2] = np.uint64(37 << 54); lwe[:,-1] = np.uint64(3<<54)
run(lwe, "single mask, 4 cts (G=4)")
lwe = rng.integers(0, 2**64, (1, o.n+1), dtype=np.uint64)
run(lwe, "random 1 ct (G=1)")
lwe = rng.integers(0, 2**64, (4, o.n+1), dtype=np.uint64)
run(lwe, "random 4 ct (G=4)")
lwe = rng.integers(0, 2**64, (16, o.n+1), dtype=np.uint64)
run(lwe, "random 16 ct (G=8)")
# c: VP single ggsw
l = o.encrypt_lwe_small(np.array([1<<63], dtype=np.uint64))
ggsw = np.stack([o.circuit_bootstrap_boolean(l[0])])
lutp = rng.integers(0, 2, (1, 1, 512)).astype(np.uint64) << np.uint64(63)
got = e.vertical_packing(lutp, ggsw); ref = o.vertical_packing(lutp[0], ggsw)
print("vp 1 ggsw: 2^%.1f" % np.log2(max(torus_absdiff(got[0], ref),1)))
