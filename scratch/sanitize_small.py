"""Small PARAM_OPT invocations of every new kernel path for compute-sanitizer (memcheck / racecheck): cluster PBS (5 ciphertexts),
warp-specialised PBS with tensor-memory parking (4 ciphertexts, forced schedule 2, G = 1 and a ragged G = 3 wave is too slow under the
sanitizer), one many_sbox of one byte (keyswitch, PBS, PFKS, Fourier, vertical packing, tcgen05 kernels)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
eng = pkg.Engine(pkg.param_opt())
eng.client_keygen(5)
data = bytes([0x53])
ct = eng.client_encrypt_bytes(data, seed=7)
out = eng.many_sbox(ct, False)          # 8 bits: cluster PBS kernel (<= 74 ciphertexts)
s = pkg.SBOX[0x53]
assert eng.client_decrypt_bytes(out[0]) == bytes([s, pkg.mul2(s), pkg.mul3(s)])
eng.set_pbs_schedule(2)                 # the same through the warp-specialised kernel (G = 1)
out = eng.sbox(ct, False)
assert eng.client_decrypt_bytes(out) == bytes([s])
eng.set_pbs_schedule(0)
print("sanitize_small ok")
