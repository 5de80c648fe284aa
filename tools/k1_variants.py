#!/usr/bin/env python3
"""Compiles nb_force.cu with experiment knobs (-DNB_EXP_*=..) to a cubin and prints the issue-model cost
of K1's hot loops (tools/sass_model.py) — no GPU needed.  Use it to rank by instruction COUNTS: round 2 showed
that the model's cycle figure is off by up to 2 % in either direction (profiles/r2_k1_variants.txt), so a
variant is only accepted after tools/k1_hw_variants.py has timed it.

  python tools/k1_variants.py "" "-DNB_EXP_KREG=0" "-DNB_EXP_KZ_UNI=1 -DNB_EXP_UNR4=2" ...
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_model  # noqa: E402


def build(flags, out):
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-fmad=false", "-cubin", "-o", out, os.path.join(ROOT, "nbodygo_b200", "csrc", "nb_force.cu")] + flags.split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stderr)


def main():
    variants = sys.argv[1:] or [""]
    for v in variants:
        with tempfile.TemporaryDirectory() as d:
            out = os.path.join(d, "k.cubin")
            try:
                build(v, out)
            except RuntimeError as e:
                print(f"[{v}] compile error: {str(e)[-300:]}")
                continue
            res = subprocess.run(["cuobjdump", "-res-usage", out], capture_output=True, text=True).stdout
            regs = {}
            name = None
            for line in res.splitlines():
                if "Function" in line:
                    name = line.split()[1].rstrip(":")
                elif "REG:" in line and name:
                    regs[name] = line.strip().split()[0] + " " + line.strip().split()[1]
            print(f"== [{v or 'baseline'}]")
            for kname, addr, n, hist, n_fp64, three, cycles in sass_model.hot_loops(out, "k_forceILi4ELi128ELi1ELi"):
                if hist.get("SEL", 0):
                    continue   # the SELF loop runs on 2 of ~3900 tiles
                if "ELi256ELi" not in kname:
                    continue
                mode = kname.split("ELi256ELi")[1][0] + " unr" + kname.split("k_forceILi4ELi128ELi1ELi")[1][0]
                others = n - n_fp64
                print(f"   mode {mode}: {cycles} cycles  ({n} instr, {n_fp64} FP64, {three} three-reg, {others} others; "
                      f"IMAD {hist.get('IMAD', 0)} MOV {hist.get('MOV', 0)} LDS {hist.get('LDS', 0)})  {regs.get(kname, '')}")


if __name__ == "__main__":
    main()
