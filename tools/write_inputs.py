#!/usr/bin/env python3
"""Writes the seeded BASELINE clouds as CSV files the reference's own loader reads
(cmd/sim/fromcsv.go:15-47; 13 columns, %.17g — the doubles survive the round trip bit for bit), and
prints the command that runs the unmodified Go server on each: the identical-input channel of
SURVEY §8(f)1 / INTEGRATION.md §5 for a box that has a Go toolchain.

  python tools/write_inputs.py --out /tmp/nbody_inputs [--configs C1,C2,C3] [--bodies N]

C4 (1,000,000 bodies, ~190 MB of CSV) is written only when asked for by name.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

COLLISION = {"C1": "elastic", "C2": "none", "C3": "elastic", "C3dense": "elastic", "C4": "elastic"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--configs", default="C1,C2,C3")
    ap.add_argument("--bodies", type=int, default=0, help="override the body count of every config")
    a = ap.parse_args()
    from nbodygo_b200 import clouds
    os.makedirs(a.out, exist_ok=True)
    for name in a.configs.split(","):
        b = clouds.config(name, n=a.bodies or None)
        path = os.path.join(a.out, f"{name.lower()}_{b.n}.csv")
        clouds.write_csv(path, b)
        back = clouds.read_csv(path)
        assert back.n == b.n and all((getattr(back, f).view("uint64") == getattr(b, f).view("uint64")).all()
                                     for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius")), "CSV round trip"
        print(f"{path}: {b.n} bodies\n  bin/server --no-render --no-barnes-hut --collision={COLLISION[name]} "
              f"--csv={path} --bodies={b.n} --threads=$(nproc) --run-millis=60000\n"
              f"  python bench.py --config {name}" + (f" --bodies {a.bodies}" if a.bodies else ""))


if __name__ == "__main__":
    main()
