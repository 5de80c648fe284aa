#!/usr/bin/env python3
"""Sum of ptxas' static stall counts over K1's hot loops (control words of the SASS) — the compiler's own
issue-cycle estimate per loop trip — next to the instruction mix.  Development tool.

  python tools/sass_stalls.py [path to .so / .cubin] [kernel substring]"""
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout.splitlines()
    funcs, cur, i = {}, None, 0
    while i < len(out):
        m = re.search(r"Function : (\S+)", out[i])
        if m:
            cur = m.group(1)
            funcs[cur] = []
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]+) \*/", out[i])
        if m and cur and i + 1 < len(out):
            m2 = re.match(r"\s+/\* 0x([0-9a-f]+) \*/", out[i + 1])
            hi = int(m2.group(1), 16) if m2 else 0
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip(), hi))
            i += 1
        i += 1
    return funcs


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nbodygo_b200", "libnbody_b200.so")
    sub = sys.argv[2] if len(sys.argv) > 2 else "k_forceILi4ELi128ELi1ELi1ELi256"
    for name, ins in functions(path).items():
        if sub not in name:
            continue
        addr = {a: k for k, (a, _, _) in enumerate(ins)}
        for k, (a, text, hi) in enumerate(ins):
            if not text.startswith("BRA") and " BRA" not in text:
                continue
            m = re.search(r"0x([0-9a-f]+)\s*$", text)
            if not m or int(m.group(1), 16) not in addr:
                continue
            t = addr[int(m.group(1), 16)]
            if not (t < k and k - t > 100):
                continue
            body = ins[t:k + 1]
            if any(op in x[1] for x in body for op in ("CALL", "LDG", "STG", "BAR", "SEL ")):
                continue
            stall = sum((h >> 41) & 0xF for _, _, h in body)
            fp64 = sum(1 for _, x, _ in body if re.match(r"(@\S+ )?(DFMA|DMUL|DADD)", x))
            yields = sum(1 for _, _, h in body if not (h >> 45) & 1)
            print(f"{name} loop @0x{ins[t][0]:04x}: {len(body)} instr, {fp64} FP64, sum of stall counts {stall}, "
                  f"{yields} yield hints")


if __name__ == "__main__":
    main()
