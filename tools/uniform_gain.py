#!/usr/bin/env python3
"""K1 with and without the uniform-mass pass (NB_UNIFORM_TILES=1/0) on an equal-mass cloud, and
on the same cloud with one mass per tile perturbed (every chunk mixed): times, force agreement.

  python tools/uniform_gain.py --n 200000 [--reps 3]

Development tool; the numbers that count come from bench.py."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def k1(capi, b, uniform, reps):
    os.environ["NB_UNIFORM_TILES"] = str(uniform)
    sim = capi.Sim(b.n)
    sim.upload(b)
    opts = capi.STEP_NO_INTEGRATE | capi.STEP_PHASE_TIMINGS | capi.STEP_COLLISIONS
    sim.step(1e-9, 1.0, opts)
    ms = min(sim.step(1e-9, 1.0, opts).ms_force for _ in range(reps))
    f = np.stack(sim.forces(), axis=1)
    pairs = sim.pairs()
    launches = sim.launch_count()
    sim.close()
    os.environ.pop("NB_UNIFORM_TILES")
    return ms, f, pairs, launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from nbodygo_b200 import capi, clouds
    peak, _ = capi.measure_fp64_peak(0, 4096)
    b = clouds.config("C4", n=a.n)
    mixed = b.copy()
    mixed.mass[::256] *= 1.0 + 2.0 ** -40   # one body per tile: no uniform tile is left
    for name, cloud in (("equal masses", b), ("one odd mass per tile", mixed)):
        ms0, f0, p0, _ = k1(capi, cloud, 0, a.reps)
        ms1, f1, p1, l1 = k1(capi, cloud, 1, a.reps)
        scale = np.abs(f0).max(axis=1, keepdims=True)
        dev = np.max(np.abs(f1 - f0) / scale)
        rate = lambda ms: 30 * a.n * (a.n - 1.0) / (ms * 1e-3) / 1e12
        print(f"{name}: general {ms0:.3f} ms ({rate(ms0) / peak * 100:.1f}% of {peak:.1f} TF/s)  "
              f"uniform pass {ms1:.3f} ms ({rate(ms1) / peak * 100:.1f}%)  gain {100 * (ms0 / ms1 - 1):.1f}%  "
              f"max |dF|/|F|inf = {dev:.2e}  pairs equal: {np.array_equal(p0, p1)}  bits equal: "
              f"{np.array_equal(f0.view(np.uint64), f1.view(np.uint64))}", flush=True)


if __name__ == "__main__":
    main()
