#!/usr/bin/env python3
"""Kernel micro-benchmark: times K1 (k_force) alone for several launch shapes.

  python tools/kbench.py --n 200000 --codes 4,2,3,12,212,404 [--collisions]

Codes are NB_FORCE_R values (see launch_force in nb_force.cu). Development tool; the
numbers that count come from bench.py."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--codes", default="4,2,1")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--collisions", action="store_true")
    a = ap.parse_args()
    from nbodygo_b200 import capi, clouds
    b = clouds.config("C4", n=a.n)
    os.environ.setdefault("NB_UNIFORM_TILES", "1")
    peak, _ = capi.measure_fp64_peak(0, 4096)
    print(f"n={a.n} fp64 peak measured {peak:.2f} TFLOP/s")
    ref = None
    for code in a.codes.split(","):
        os.environ["NB_FORCE_R"] = code
        sim = capi.Sim(b.n)
        sim.upload(b)
        opts = capi.STEP_NO_INTEGRATE | capi.STEP_PHASE_TIMINGS | (capi.STEP_COLLISIONS if a.collisions else 0)
        sim.step(1e-9, 1.0, opts)
        ms = [sim.step(1e-9, 1.0, opts).ms_force for _ in range(a.reps)]
        fx, _, _ = sim.forces()
        same = "" if ref is None else f" bits_equal={np.array_equal(fx.view(np.uint64), ref.view(np.uint64))}"
        ref = fx if ref is None else ref
        best = min(ms)
        rate = b.n * (b.n - 1.0) / (best * 1e-3)
        print(f"code {code:>4}: {best:9.3f} ms  {rate:.4e} pairs/s  {30 * rate / 1e12:6.2f} TFLOP/s "
              f"({30 * rate / 1e12 / peak * 100:5.1f}% of measured peak){same}", flush=True)
        sim.close()


if __name__ == "__main__":
    main()
