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

# 256-body tiles, the uniform-mass / per-body-mass split, deleted bodies in the array (dead_j_sweep), the range
# downloads, nb_set_forces and the sharded-upload entry point on a single handle
n2 = 20_000
c = clouds.uniform_cube(n2, 200.0, 1.0, 1e12, vmax=100.0, seed=6)
c.radius[:] = rng.uniform(0.5, 7.0, n2)
c.behavior[rng.random(n2) < 0.3] = SUBSUME
dead = rng.random(n2) < 0.1
c.flags[dead] &= ~np.uint8(1)
sim = capi.Sim(n2)
sim.upload_shard(n2, 0, n2, c.x, c.y, c.z, c.vx, c.vy, c.vz, c.mass, c.radius, behavior=c.behavior, flags=c.flags)
for _ in range(2):
    r = sim.step(1e-4, 0.9)
fx, fy, fz = sim.forces()
sim.set_forces(100, 50, fx[100:150], fy[100:150], fz[100:150])
xs = np.zeros(500)
sim.download_range_into(1000, 500, x=xs)
xyz, ex = np.zeros((500, 3), dtype=np.float32), np.zeros(500, dtype=np.uint8)
sim.render_range(1000, 500, xyz, ex)
print("sanitize_mixed (large tiles, dead bodies) ok: subsumed", r.n_subsumed, "dead", r.n_dead, "events", r.n_host_events)
sim.close()
