"""Single-kernel runs for ncu: PBS (one wave, G=3), PFKS, KS, VP at PARAM_OPT shapes."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
eng = pkg.Engine(pkg.param_opt(), device=0, stream=stream.cuda_stream)
eng.client_keygen(1)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 444
lwe = torch.randint(-2**62, 2**62, (count * (eng.n + 1),), dtype=torch.int64, device=dev)
lut = torch.full((512,), -(1 << 48), dtype=torch.int64, device=dev)
out = torch.zeros(count * eng.lw, dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    e0.record(stream)
    eng.bootstrap_dev(lwe.data_ptr(), count, lut.data_ptr(), 1 << 62, 1 << 48, out.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"pbs count={count}: {ms:.3f} ms -> {count/ms*1e3:.0f} PBS/s, {count*407608320/ms*1e-9:.2f} TFLOP/s")
# one many_sbox batch of 16 bytes x (count/128) blocks for the other stages
nblk = max(1, count // 128)
data = bytes(np.random.default_rng(0).integers(0, 256, 16 * nblk, dtype=np.uint8))
ct = eng.client_encrypt_bytes(data, seed=3)
eng.profile(True)
res = eng.many_sbox(ct, False)
print("stages (ms, groups):", eng.profile_report())
dec = eng.client_decrypt_bytes(res[:, 0])
assert dec == bytes(pkg.SBOX[b] for b in data)
print("fp64 peak", eng.measure_fp64_peak())
