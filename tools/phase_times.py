#!/usr/bin/env python3
"""Per-phase device time of a cycle (K0 / K1 / K3 / K4) for the C1 (Sim3 geometry) collision sim and
a few cloud sizes: averages over the cycles of a short run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodygo_b200 import capi, clouds  # noqa: E402


def run(name, b, ts, steps=400):
    sim = capi.Sim(b.n)
    sim.upload(b)
    acc = np.zeros(8)
    mx_rounds = 0
    for k in range(steps):
        r = sim.step(ts, 1.0, capi.STEP_DEFAULT | capi.STEP_PHASE_TIMINGS)
        acc += (r.ms_prep, r.ms_force, r.ms_exchange, r.ms_resolve, r.ms_integrate, r.ms_total, r.n_pairs, r.resolve_rounds)
        mx_rounds = max(mx_rounds, r.resolve_rounds)
    acc /= steps
    print(f"{name:28s} n={b.n:7d}: K0 {acc[0]*1e3:6.1f}  K1 {acc[1]*1e3:8.1f}  K3 {acc[3]*1e3:7.1f}  K4 {acc[4]*1e3:6.1f}  "
          f"total {acc[5]*1e3:8.1f} us | pairs/cycle {acc[6]:8.1f} rounds avg {acc[7]:5.1f} max {mx_rounds}", flush=True)
    sim.close()


run("C1 Sim3 clusters", clouds.config("C1", n=1001), 1e-9)
run("C1 Sim3 clusters 3001", clouds.config("C1", n=3001), 1e-9)
run("C3 cube 100k", clouds.config("C3"), 1e-9, steps=20)
run("C3-dense cube 100k (r x4)", clouds.config("C3dense"), 1e-9, steps=20)
