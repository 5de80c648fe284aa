#!/usr/bin/env python3
"""Static cost of K1's hot loops from the SASS of the built library (no GPU needed).

  python tools/sass_model.py [--kernel SUBSTRING] [--so PATH]

For every loop of the selected k_force instantiations that contains more than 40 instructions, at
least 30 of them FP64, and no call, barrier or global / local memory access, prints: instruction count by opcode, FP64 instructions, distinct 64-bit register operands
they read (the `.reuse` cache of the previous instruction taken into account), and the cycles per
iteration predicted by the issue model measured on B200 (profiles/r1_summary.md):

    cycles = sum over FP64 instructions of max(2, distinct register operands read)
           + 1 for every other instruction

One iteration of the production loop (R = 4 i-bodies x 2 j-bodies) is 8 pairs.  The model reproduced
the measured launch of the general pass (291 predicted, 288.4 measured cycles per 8 pairs) and the
gain of the uniform-mass pass (274 predicted: -5.8 %; measured -5.3 % with 1 of 32 chunks general).
"""
import argparse
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP64 = ("DFMA", "DMUL", "DADD")


def sass_functions(so):
    """{mangled name: [(address, text), ...]} for every k_force instantiation in the library."""
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1) if "k_force" in m.group(1) else None
            if cur:
                funcs[cur] = []
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
        if m:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def opcode(text):
    f = text.split()
    op = f[1] if f[0].startswith("@") else f[0]
    return op.split(".")[0]


def loops(instrs, min_len=40):
    """Innermost backward-branch loops [(start index, end index)] with more than min_len instructions."""
    addr = {a: i for i, (a, _) in enumerate(instrs)}
    found = []
    for i, (_, text) in enumerate(instrs):
        if opcode(text) != "BRA":
            continue
        m = re.search(r"0x([0-9a-f]+)\s*$", text)
        if not m:
            continue
        t = addr.get(int(m.group(1), 16))
        if t is not None and t < i and i - t > min_len:
            found.append((t, i))
    # innermost only: drop loops that contain another found loop
    return [(a, b) for (a, b) in found if not any((a <= c and d <= b) and (a, b) != (c, d) for (c, d) in found)]


def cost(body):
    """(histogram, n_fp64, reads, three_read, predicted cycles, has_call_or_local)."""
    hist, cache = {}, {}
    n_fp64 = reads = three = cycles = 0
    impure = False
    for _, text in body:
        op = opcode(text)
        hist[op] = hist.get(op, 0) + 1
        if op in ("CALL", "LDL", "STL", "LDG", "STG", "BAR", "BSSY"):
            impure = True
        if op in FP64:
            f = text.split()
            ops = [o.strip() for o in text[text.index(f[1] if f[0].startswith("@") else f[0]) + len(
                f[1] if f[0].startswith("@") else f[0]):].split(",")][1:]
            rd, new = set(), {}
            for slot, o in enumerate(ops):
                m = re.match(r"-?\|?(R\d+)\|?(\.reuse)?$", o)
                if not m:
                    continue
                if cache.get(slot) != m.group(1):
                    rd.add(m.group(1))
                if m.group(2):
                    new[slot] = m.group(1)
            cache = new
            n_fp64 += 1
            reads += len(rd)
            three += len(rd) >= 3
            cycles += max(2, len(rd))
        else:
            cache = {}
            cycles += 1
    return hist, n_fp64, reads, three, cycles, impure


def hot_loops(so, kernel_filter=""):
    """[(kernel, start address, n_instr, hist, n_fp64, three_read, cycles)] of the pure FP64 loops."""
    rows = []
    for name, instrs in sass_functions(so).items():
        if kernel_filter and kernel_filter not in name:
            continue
        for a, b in loops(instrs):
            body = instrs[a:b + 1]
            hist, n_fp64, reads, three, cycles, impure = cost(body)
            if impure or n_fp64 < 30:
                continue   # the redo / exact paths, not the fast pass
            rows.append((name, instrs[a][0], len(body), hist, n_fp64, three, cycles))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--so", default=os.path.join(ROOT, "nbodygo_b200", "libnbody_b200.so"))
    ap.add_argument("--kernel", default="k_forceILi4ELi128ELi1ELi1ELi256", help="substring of the mangled name")
    a = ap.parse_args()
    for name, addr, n, hist, n_fp64, three, cycles in hot_loops(a.so, a.kernel):
        kind = "SELF" if hist.get("SEL", 0) else "    "
        others = n - n_fp64
        print(f"{name}\n  loop @0x{addr:04x} {kind}: {n} instructions, {n_fp64} FP64 ({three} read three registers), "
              f"{others} others -> {cycles} cycles per iteration (model)\n    " +
              " ".join(f"{k}:{v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    sys.exit(main())
