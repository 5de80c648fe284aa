#!/usr/bin/env python3
"""Small mixed-behaviour cycles (elastic, subsume chains, fragment decisions, dead bodies, compaction,
append) for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_mixed.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodygo_b200 import capi, clouds  # noqa: E402
from nbodygo_b200.bodies import FRAGMENT, NONE, SUBSUME  # noqa: E402

rng = np.random.default_rng(9)
n = 1500
b = clouds.uniform_cube(n, 70.0, 1.5, 1e12, vmax=300.0, seed=4)
b.radius[:] = rng.uniform(0.4, 6.0, n)
b.behavior[rng.random(n) < 0.25] = SUBSUME
b.behavior[rng.random(n) < 0.15] = FRAGMENT
b.behavior[rng.random(n) < 0.05] = NONE
b.frag_factor[:] = 0.05
b.frag_step[:] = 100.0
sim = capi.Sim(n + 64)
sim.upload(b)
tot = dict(pairs=0, sub=0, hev=0)
for k in range(4):
    r = sim.step(1e-4, 0.9)
    tot["pairs"] += r.n_pairs; tot["sub"] += r.n_subsumed; tot["hev"] += r.n_host_events
    sim.pairs(); sim.host_events(); sim.forces()
    if k == 1:
        sim.compact()
        sim.append(clouds.uniform_cube(7, 10.0, 1.0, 1e12, seed=5), R=0.9)
r = sim.step(1e-4, 0.9, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE)
sim.host_events()
sim.download()
print("sanitize_mixed ok:", tot, "n =", sim.count())
sim.close()
