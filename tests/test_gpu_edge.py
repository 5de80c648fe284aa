"""Edge cases of the state the host can inject through mod-body (globals.SafeParseFloat accepts
"NaN" and "Inf", cmd/globals/globals.go:97-102): NaN radius, non-finite position, zero mass."""
import numpy as np
import pytest

from nbodygo_b200 import clouds
from nbodygo_b200.bodies import F_EXISTS

pytestmark = pytest.mark.gpu


def _both(b, ts=1e-3, R=1.0):
    from nbodygo_b200 import capi
    from oracle.oracle import OracleSim
    o = OracleSim(b.copy())
    o.compute()
    pairs = o.collision_pairs()
    ex, ey, ez, fn = o.compute_exact()
    o.process_mods()
    o.update(ts, R)
    sim = capi.Sim(b.n)
    sim.upload(b)
    res = sim.step(ts, R)
    out = dict(res=res, pairs=sim.pairs(), forces=sim.forces(), state=sim.download(), o=o, ref_pairs=pairs,
               exact=(ex, ey, ez, fn))
    sim.close()
    return out


def test_nan_radius_body_matches_reference_predicate():
    # radius = NaN: dist > r_i + NaN and dist <= r_i + NaN are both false ⇒ the body neither exerts
    # nor feels gravity and never collides (cmd/body/body.go:199-204,219) — for everybody in its tile too
    b = clouds.uniform_cube(700, 60.0, 2.0, 1e13, vmax=5.0, seed=61)
    b.radius[[5, 300]] = np.nan
    r = _both(b)
    assert np.array_equal(r["pairs"], r["ref_pairs"])
    assert not any(5 in p or 300 in p for p in r["pairs"].tolist())
    fx, fy, fz = r["forces"]
    ex, ey, ez, fn = r["exact"]
    assert fx[5] == 0 and fx[300] == 0 and ex[5] == 0
    err = np.max(np.abs(np.stack([fx - ex, fy - ey, fz - ez])), axis=0)
    assert np.all(err <= 1e-12 * np.where(fn > 0, fn, 1.0))
    assert np.allclose(r["state"].x, r["o"].b.x, rtol=1e-12, atol=1e-12)


def test_zero_mass_live_body():
    # a live body of mass 0 exerts no force; Update divides by its mass: 0*f/0 = NaN ⇒ NaN cull
    b = clouds.uniform_cube(300, 80.0, 1.0, 1e13, vmax=5.0, seed=62)
    b.mass[7] = 0.0
    r = _both(b)
    assert np.array_equal(r["state"].exists, r["o"].b.exists)
    assert not r["state"].exists[7] and r["res"].n_culled == 1
    assert np.array_equal(r["pairs"], r["ref_pairs"])


def test_non_finite_position_is_contained():
    # Documented deviation: in the reference one body at x = Inf turns every force into
    # NaN (0 * Inf) and the whole simulation is culled; here the body is inert as a j-body, receives
    # NaN itself and is the only one removed.  The step must neither hang nor poison the others.
    b = clouds.uniform_cube(500, 80.0, 1.0, 1e13, vmax=5.0, seed=63)
    b.x[11] = np.inf
    b.y[12] = 1e200
    from nbodygo_b200 import capi
    sim = capi.Sim(b.n)
    sim.upload(b)
    res = sim.step(1e-3, 1.0)
    g = sim.download()
    dead = np.where(~g.exists)[0].tolist()
    assert dead == [11, 12] and res.n_culled == 2
    others = np.ones(b.n, dtype=bool)
    others[[11, 12]] = False
    assert np.isfinite(g.x[others]).all() and np.isfinite(g.vx[others]).all()
    # and the others moved exactly as if the two bodies did not exist
    b2 = b.copy()
    b2.flags[[11, 12]] = 0
    b2.x[11] = 0.0
    b2.y[12] = 0.0
    sim2 = capi.Sim(b.n)
    sim2.upload(b2)
    sim2.step(1e-3, 1.0)
    g2 = sim2.download()
    assert np.array_equal(g.x[others].view(np.uint64), g2.x[others].view(np.uint64))
    assert np.array_equal(g.vz[others].view(np.uint64), g2.vz[others].view(np.uint64))
    sim.close()
    sim2.close()
