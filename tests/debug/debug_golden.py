"""Development diagnostic: per-step deviation of the device from a golden scene."""
import sys
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import load_golden, scene_bodies, scene_step_arrays, unhex
from nbodygo_b200 import capi

name = sys.argv[1] if len(sys.argv) > 1 else "dense_mixed_48"
scene = [s for s in load_golden() if s["name"] == name][0]
b = scene_bodies(scene)
ts, R = unhex(scene["ts"]), unhex(scene["R"])
sim = capi.Sim(b.n)
sim.upload(b)
live = np.array([r["exists"] for r in scene["init"]])
for k, step in enumerate(scene["steps"]):
    exp = scene_step_arrays(step)
    res = sim.step(ts, R)
    fx, fy, fz = sim.forces()
    got = np.stack([fx, fy, fz], axis=1)
    d = np.abs(got - exp["forces"])
    scale = np.max(np.abs(exp["forces"][live]))
    i = np.unravel_index(np.argmax(np.where(live[:, None], d, 0)), d.shape)
    print(f"step {k}: max|dF|={d[live].max():.3e} scale={scale:.3e} worst body {i} got={got[i]:.17g} exp={exp['forces'][i]:.17g}"
          f" pairs={res.n_pairs} hev={res.n_host_events} subsumed={res.n_subsumed} dead={res.n_dead}")
    g = sim.download()
    print("   mass equal:", np.array_equal(g.mass, exp["mass"]), "exists equal:", np.array_equal(g.exists, exp["exists"]),
          "max|dx|", np.nanmax(np.abs(g.x - exp["x"])), "max|dvx|", np.nanmax(np.abs(g.vx - exp["vx"])))
    bad = np.where(g.mass != exp["mass"])[0]
    if len(bad):
        print("   mass mismatch at", bad, g.mass[bad], exp["mass"][bad])
    live = exp["exists"]
sim.close()
