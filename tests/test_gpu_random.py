"""Randomised parity sweep: small clouds with random radii, behaviours, dead / fragmenting bodies
and restitution, several cycles each, GPU (through the C ABI) against the oracle."""
import numpy as np
import pytest

from nbodygo_b200.bodies import ELASTIC, F_EXISTS, F_FRAGMENTING, FRAGMENT, NONE, SUBSUME, BodyArrays

pytestmark = pytest.mark.gpu


def random_cloud(rng):
    n = int(rng.integers(2, 700))
    side = float(rng.uniform(8, 60)) * n ** (1 / 3)
    b = BodyArrays.from_fields(*(rng.uniform(-side / 2, side / 2, n) for _ in range(3)),
                               *(rng.uniform(-30, 30, n) for _ in range(3)),
                               rng.uniform(1e9, 1e13, n), rng.uniform(0.2, 4.0, n) * rng.choice([1, 1, 3]))
    b.behavior[:] = rng.choice([NONE, SUBSUME, ELASTIC, ELASTIC, ELASTIC, FRAGMENT], n).astype(np.uint8)
    b.frag_factor[:] = rng.uniform(0.0, 5.0, n)
    b.frag_step[:] = rng.uniform(1.0, 300.0, n)
    dead = rng.random(n) < 0.03
    b.flags[dead] = 0
    b.mass[dead] = 0
    b.flags[(rng.random(n) < 0.02) & ~dead] |= F_FRAGMENTING
    b.rest[:] = rng.choice([1.0, 0.8, 0.3])
    return b


@pytest.mark.parametrize("seed", range(24))
def test_random_cloud_cycles(seed):
    from nbodygo_b200 import capi
    from oracle.oracle import EV_FRAGMENT, EV_SUBSUME, OracleSim
    rng = np.random.default_rng(1000 + seed)
    b = random_cloud(rng)
    R, ts = float(rng.choice([1.0, 0.7])), float(rng.choice([1e-3, 1e-5]))
    o = OracleSim(b.copy())
    sim = capi.Sim(b.n)
    sim.upload(b)
    last_fn = np.ones(b.n)
    for step in range(3):
        computing = ((o.b.flags & F_EXISTS) != 0) & ((o.b.flags & F_FRAGMENTING) == 0)
        o.compute()
        ref_pairs = o.collision_pairs()
        ref_sub = sorted((int(e["a"]), int(e["b"]), float(e["dist"])) for e in o.events if e["kind"] == EV_SUBSUME)
        ex, ey, ez, fn = o.compute_exact()
        # the whole event queue — elastic, fragment decisions, subsumes — is resolved on the device in
        # the reference's serial order, so the oracle simply runs ProcessMods
        o.process_mods()
        ref_frag = sorted((int(e["a"]), int(e["b"])) for e in o.host_events if e["kind"] == EV_FRAGMENT)
        o.update(ts, R)
        res = sim.step(ts, R)
        assert np.array_equal(sim.pairs(), ref_pairs), f"seed {seed} step {step}: pair set"
        hev = sim.host_events()
        assert [(int(e["a"]), int(e["b"]), float(e["dist"])) for e in hev if e["kind"] == capi.EV_SUBSUME] == ref_sub
        assert sorted((int(e["a"]), int(e["b"])) for e in hev if e["kind"] == capi.EV_FRAGMENT) == ref_frag
        fx, fy, fz = sim.forces()
        # bodies that do not compute this cycle (dead / fragmenting) keep their previous force
        rx, ry, rz = (np.where(computing, e, f) for e, f in ((ex, o.fx), (ey, o.fy), (ez, o.fz)))
        last_fn = np.where(computing, fn, last_fn)
        err = np.max(np.abs(np.stack([fx - rx, fy - ry, fz - rz])), axis=0)
        assert np.all(err <= 1e-12 * np.where(last_fn > 0, last_fn, 1.0)), f"seed {seed} step {step}: forces"
        g = sim.download()
        assert np.array_equal(g.exists, o.b.exists)
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            a, r = getattr(g, f), getattr(o.b, f)
            m = ~np.isnan(r)
            assert np.array_equal(np.isnan(a), np.isnan(r))
            assert np.allclose(a[m], r[m], rtol=1e-10, atol=1e-10 * max(np.max(np.abs(r[m]), initial=0.0), 1e-300)), \
                f"seed {seed} step {step}: {f}"
        # masses (ResolveSubsume) and the flags of initiateFragmentation / SetNotExists: bit-exact, no
        # host patch between cycles
        assert np.array_equal(g.mass, o.b.mass), f"seed {seed} step {step}: mass"
        assert np.array_equal(g.flags & (F_EXISTS | F_FRAGMENTING), o.b.flags & (F_EXISTS | F_FRAGMENTING))
        assert np.array_equal(g.behavior, o.b.behavior)
    sim.close()
