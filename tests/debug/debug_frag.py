"""Debug: fragments with identical velocities overlapping a target — where do NaNs come from?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from nbodygo_b200.bodies import BodyArrays, ELASTIC, F_EXISTS
from oracle.oracle import OracleSim

rng = np.random.default_rng(1)
n = 250
p = rng.uniform(-1, 1, size=(4000, 3)); p = p[(p * p).sum(1) <= 1][: n - 1] * 9 + np.array([1015.0, 0, 0])
x = np.concatenate([[1000.0], p[:, 0]]); y = np.concatenate([[0.0], p[:, 1]]); z = np.concatenate([[0.0], p[:, 2]])
vx = np.concatenate([[-0.9e9], np.full(n - 1, -1e9)]); vy = np.concatenate([[0.0], np.full(n - 1, 2e8)]); vz = np.zeros(n)
m = np.concatenate([[1e12], np.full(n - 1, 1e12 / 249)]); r = np.concatenate([[10.0], np.ones(n - 1)])
b = BodyArrays.from_fields(x, y, z, vx, vy, vz, m, r)
o = OracleSim(b.copy())
use_gpu = len(sys.argv) > 1 and sys.argv[1] == "gpu"
if use_gpu:
    from nbodygo_b200 import capi
    sim = capi.Sim(n); sim.upload(b)
for step in range(6):
    o.compute(); pairs = len(o.collision_pairs()); o.process_mods()
    nanv = np.isnan(o.b.vx).sum()
    o.update(1e-12, 1.0)
    msg = f"step {step}: oracle pairs={pairs} nan_v_after_resolve={nanv} dead={(~o.b.exists).sum()}"
    if use_gpu:
        res = sim.step(1e-12, 1.0); g = sim.download()
        msg += f" | gpu pairs={res.n_pairs} culled={res.n_culled} dead={res.n_dead} rounds={res.resolve_rounds}"
        if res.n_culled:
            bad = np.where(~g.exists)[0][:5]; msg += f" first dead {bad}"
    print(msg)
