"""Decode the control word of sm_100 SASS (cuobjdump -sass output) and apply the single-warp issue model of
/opt/skills/guides/B300_MICROARCH.md to a range of instructions: prints per-instruction stall / wbar / rbar /
wait_mask and an estimate of the isolated-warp cycles of the range (LDS 29, FP64 by measurement parameter).
usage: python scratch/sass_ctl.py file.sass START_ADDR END_ADDR [-v]"""
import re, sys

def parse(path):
    ins = []
    lines = open(path).read().split("\n")
    i = 0
    pat = re.compile(r"^\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/")
    pat2 = re.compile(r"^\s*/\* 0x([0-9a-f]{16}) \*/")
    while i < len(lines):
        m = pat.match(lines[i])
        if m and i + 1 < len(lines):
            m2 = pat2.match(lines[i + 1])
            if m2:
                lo, hi = int(m.group(3), 16), int(m2.group(1), 16)
                w = (hi << 64) | lo
                ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(),
                                stall=(w >> 105) & 0xF, yld=(w >> 109) & 1, wbar=(w >> 110) & 7,
                                rbar=(w >> 113) & 7, wait=(w >> 116) & 0x3F))
                i += 2
                continue
        i += 1
    return ins

LAT = [("LDS", 30), ("LDG", 400), ("LDTM", 30), ("DFMA", 10), ("DADD", 10), ("DMUL", 10), ("SYNCS", 90), ("MUFU", 20), ("STS", 10), ("SHFL", 25)]
def lat_of(t):
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    for k, v in LAT:
        if op.startswith(k):
            return v
    return 15

def model(ins, a0, a1, verbose=False):
    T = 0
    sb = [0] * 6
    n = nf = 0
    for x in ins:
        if x["addr"] < a0 or x["addr"] >= a1:
            continue
        arm = max([sb[s] for s in range(6) if x["wait"] >> s & 1] + [0])
        T0 = T
        T = max(T, arm)
        if verbose:
            print(f'{x["addr"]:05x} T={T:6d} (+{T - T0:3d} sbwait) st={x["stall"]:2d} w={x["wbar"]} r={x["rbar"]} m={x["wait"]:02x}  {x["text"][:70]}')
        if x["wbar"] < 6:
            sb[x["wbar"]] = max(sb[x["wbar"]], T + lat_of(x["text"]))
        if x["rbar"] < 6:
            sb[x["rbar"]] = max(sb[x["rbar"]], T + 6)
        T += max(1, x["stall"])
        n += 1
        op = x["text"].split()[1] if x["text"].startswith("@") else x["text"].split()[0]
        if op[:4] in ("DFMA", "DADD", "DMUL"):
            nf += 1
    return T, n, nf

if __name__ == "__main__":
    ins = parse(sys.argv[1])
    a0, a1 = int(sys.argv[2], 16), int(sys.argv[3], 16)
    T, n, nf = model(ins, a0, a1, "-v" in sys.argv)
    print(f"range {a0:x}..{a1:x}: {n} instr, {nf} fp64, isolated-warp cycles ~{T}, fp64 pipe cycles {2 * nf}")

def loops(ins):
    """backward branches = loops: (start, end) address pairs"""
    out = []
    for x in ins:
        m = re.search(r"\bBRA(?:\.U)?(?:\.ANY)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", x["text"])
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= x["addr"]:
                out.append((tgt, x["addr"] + 16))
    return out
