"""Cluster PBS kernel (schedule 3) against the warp-specialised kernel (schedule 2) on real keys: phases and timing."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
eng = pkg.Engine(pkg.param_opt(), device=0, stream=stream.cuda_stream)
eng.client_keygen(5)
lwe_sk, glwe_sk = eng.client_secret_keys()
rng = np.random.default_rng(1)
for count in (1, 5, 32, 74):
    bits = rng.integers(0, 2, count)
    data = bytes(int(b) for b in bits)             # byte value 0/1: bit 0 of each byte carries the message
    ct = eng.client_encrypt_bytes(data, seed=7)[:, 0, :]          # [count][lw]: LSB block of every byte
    ks = eng.keyswitch(ct)
    ks[:, -1] += np.uint64(1 << 62)
    lut = np.full(512, (1 << 64) - (1 << 48), dtype=np.uint64)
    res = {}
    for sched in (2, 3):
        eng.set_pbs_schedule(sched)
        out = eng.bootstrap(ks, lut)
        with np.errstate(over="ignore"):
            ph = out[:, -1] - (out[:, :-1] * glwe_sk).sum(axis=1, dtype=np.uint64) + np.uint64(1 << 48)
        dec = ((ph + np.uint64(1 << 48)) >> np.uint64(49)) & np.uint64(1)
        err = (ph - (bits.astype(np.uint64) << np.uint64(49))).astype(np.int64)
        res[sched] = (ph, dec, err)
        assert np.array_equal(dec, bits.astype(np.uint64)), (sched, count, dec, bits)
    d = (res[2][0] - res[3][0]).astype(np.int64)
    print(f"count {count}: bits ok; max |phase(ws) - phase(cl2)| = 2^{np.log2(np.abs(d).max() + 1):.1f}; err std ws 2^{np.log2(res[2][2].std() + 1):.1f} cl2 2^{np.log2(res[3][2].std() + 1):.1f}")
    # device-resident timing
    lwe = torch.from_numpy(ks.view(np.int64)).to(dev).contiguous()
    lutd = torch.from_numpy(lut.view(np.int64)).to(dev)
    outd = torch.zeros(count * eng.lw, dtype=torch.int64, device=dev)
    for sched in (2, 3):
        eng.set_pbs_schedule(sched)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.bootstrap_dev(lwe.data_ptr(), count, lutd.data_ptr(), 0, 0, outd.data_ptr())
        e0.record(stream)
        for _ in range(3):
            eng.bootstrap_dev(lwe.data_ptr(), count, lutd.data_ptr(), 0, 0, outd.data_ptr())
        e1.record(stream); torch.cuda.synchronize()
        print(f"   schedule {sched}: {e0.elapsed_time(e1) / 3:.3f} ms per launch of {count}")
eng.set_pbs_schedule(0)
print("ok")
