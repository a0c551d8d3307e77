"""List the loops of a kernel's SASS with the single-warp cycle estimate of one pass through each body.
usage: python scratch/sass_loops.py obj.o mangled_kernel_name"""
import subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_ctl
obj, fun = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
open("/tmp/sass/_cur.sass", "w").write(sass)
ins = sass_ctl.parse("/tmp/sass/_cur.sass")
for a0, a1 in sorted(set(sass_ctl.loops(ins))):
    if a1 - a0 < 0x200: continue
    T, n, nf = sass_ctl.model(ins, a0, a1)
    print(f"loop {a0:05x}..{a1:05x}: {n:5d} instr {nf:4d} fp64  isolated ~{T:5d} cycles  (fp64 pipe {2*nf})")
