"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bars (north star / SURVEY §7 H3,H5):
  * collision-pair sets and subsume events: bit-exact (same set, same order after sorting)
  * forces: |F_gpu - F_exact|_inf <= 1e-12 * sum_j |f_ij|  (normwise; summation order differs)
  * post-collision velocities: 1e-12 relative to the pair's speed scale (libm last-ulp differences)
  * integration: bit-exact given identical force/velocity inputs, else the force tolerance propagated
"""
import os

import numpy as np
import pytest

from helpers import load_golden, scene_bodies, scene_step_arrays, unhex
from nbodygo_b200 import clouds
from nbodygo_b200.bodies import (ELASTIC, F_COLLIDED, F_EXISTS, F_FRAGMENTING, FRAGMENT, NONE, SUBSUME,
                                 BodyArrays)

pytestmark = pytest.mark.gpu

FORCE_TOL = 1e-12  # normwise, stated by BASELINE.json north_star


@pytest.fixture(scope="module")
def capi():
    from nbodygo_b200 import capi as c
    c.load()
    return c


def oracle_sim(b):
    from oracle.oracle import OracleSim
    return OracleSim(b)


def assert_forces(fx, fy, fz, o, tol=FORCE_TOL):
    ex, ey, ez, fn = o.compute_exact()
    err = np.max(np.abs(np.stack([fx - ex, fy - ey, fz - ez])), axis=0)
    scale = np.where(fn > 0, fn, 1.0)
    worst = np.max(err / scale)
    assert worst <= tol, f"normwise force error {worst:.3e} > {tol}"
    return worst


def run_both(capi, b, ts, R, opts=None, steps=1):
    """Steps the oracle and the GPU side by side; yields per-step artefacts."""
    o = oracle_sim(b.copy())
    sim = capi.Sim(max(b.n, 1))
    sim.upload(b)
    out = []
    for _ in range(steps):
        o.compute()
        ref_pairs = o.collision_pairs()
        ref_events = o.events.copy()
        ref_forces = (o.fx.copy(), o.fy.copy(), o.fz.copy())
        o.process_mods()
        o.update(ts, R)
        res = sim.step(ts, R, capi.STEP_DEFAULT if opts is None else opts)
        out.append(dict(res=res, pairs=sim.pairs(), ref_pairs=ref_pairs, ref_events=ref_events,
                        forces=sim.forces(), ref_forces=ref_forces, hev=sim.host_events(),
                        state=sim.download(), ref=o.b.copy(), render=sim.render(),
                        ref_render=(o.render_xyz.copy(), o.render_exists.copy())))
    sim.close()
    return out


# ---------------------------------------------------------------- golden scenes
@pytest.mark.parametrize("scene", load_golden(), ids=lambda s: s["name"])
def test_golden_scene(capi, scene):
    b = scene_bodies(scene)
    ts, R = unhex(scene["ts"]), unhex(scene["R"])
    sim = capi.Sim(b.n)
    sim.upload(b)
    for k, step in enumerate(scene["steps"]):
        exp = scene_step_arrays(step)
        res = sim.step(ts, R)
        live = np.array([r["exists"] for r in scene["init"]]) if k == 0 else prev_exists
        fx, fy, fz = sim.forces()
        got_f = np.stack([fx, fy, fz], axis=1)
        scale = np.max(np.abs(exp["forces"][live])) if live.any() else 1.0
        assert np.allclose(got_f[live], exp["forces"][live], rtol=1e-12, atol=1e-13 * scale)
        exp_pairs = [(a, b_) for kind, a, b_, _ in step["events"] if kind == "collision"]
        assert [tuple(p) for p in sim.pairs()] == sorted(exp_pairs)
        exp_sub = sorted((a, b_, unhex(d)) for kind, a, b_, d in step["events"] if kind == "subsume")
        hev = sim.host_events()
        assert [(int(e["a"]), int(e["b"]), float(e["dist"])) for e in hev if e["kind"] == capi.EV_SUBSUME] == exp_sub
        got = sim.download()
        # ResolveSubsume runs on the device, in event order: masses and Exists are part of the state
        assert np.array_equal(got.mass, exp["mass"]), "mass (ResolveSubsume) must be bit-exact"
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            a, r = getattr(got, f), exp[f]
            m = ~np.isnan(r)
            assert np.array_equal(np.isnan(a), np.isnan(r)), f
            s = max(np.max(np.abs(r[m])), 1e-300) if m.any() else 1.0
            assert np.allclose(a[m], r[m], rtol=1e-11, atol=1e-13 * s), f
        assert np.array_equal(got.exists, exp["exists"])
        prev_exists = exp["exists"]
        assert res.n_dead == int((~exp["exists"]).sum())
    sim.close()


def test_kat1_reference_assertion(capi):
    # cmd/runner/workpool_test.go:41-56 — the reference's own check (Vx != 0), plus the derived bits
    b = BodyArrays.from_fields([1, 22], [1, 22], [1, 22], [0, 0], [0, 0], [0, 0], [1, 1], [1, 1])
    sim = capi.Sim(2)
    sim.upload(b)
    sim.step(1.0, 1.0)
    got = sim.download()
    assert got.vx[0] != 0 and got.vx[1] != 0
    fx, _, _ = sim.forces()
    assert abs(fx[0] - 2.912062242103079e-14) <= 4 * np.spacing(2.912062242103079e-14)
    assert abs(got.x[0] - 1.000000000000029) <= np.spacing(1.0)
    sim.close()


# ---------------------------------------------------------------- random clouds
@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (31, 3), (255, 4), (256, 5), (257, 6), (1000, 7), (4097, 8)])
def test_cloud_forces_and_pairs(capi, n, seed):
    b = clouds.uniform_cube(n, 60.0 * max(n, 8) ** (1 / 3), 2.0, 1e15, vmax=10.0, seed=seed)
    (s,) = run_both(capi, b, 1e-3, 1.0)
    assert np.array_equal(s["pairs"], s["ref_pairs"])
    o = oracle_sim(b.copy())
    assert_forces(*s["forces"], o)
    assert s["res"].n_pairs == len(s["ref_pairs"])


def test_collisions_off_config_c2(capi):
    # BASELINE config 2 geometry at reduced n: behaviour None ⇒ no events, overlap mask still applies
    b = clouds.config("C2", n=3000)
    (s,) = run_both(capi, b, 1e-9, 1.0, opts=0)
    assert len(s["pairs"]) == 0 and s["res"].n_pairs == 0
    assert_forces(*s["forces"], oracle_sim(b.copy()))
    for f in ("x", "vx", "y", "vy"):
        r = getattr(s["ref"], f)
        assert np.allclose(getattr(s["state"], f), r, rtol=1e-10, atol=1e-12 * np.max(np.abs(r)))


def test_dense_cloud_resolve_matches_serial_order(capi):
    # many bodies in several collisions at once: the round-based resolve must equal the
    # reference's serial reverse-arrival order
    b = clouds.uniform_cube(1500, 60.0, 2.5, 1e12, vmax=100.0, seed=21)
    (s,) = run_both(capi, b, 1e-4, 0.9)
    assert len(s["ref_pairs"]) > 1000
    assert np.array_equal(s["pairs"], s["ref_pairs"])
    assert s["res"].resolve_rounds > 2
    vscale = 100.0
    for f in ("vx", "vy", "vz"):
        assert np.allclose(getattr(s["state"], f), getattr(s["ref"], f), rtol=0, atol=1e-11 * vscale), f
    # collided bodies ignore this cycle's force (body.go:118-123) and flags are cleared afterwards
    assert not (s["state"].flags & F_COLLIDED).any()


def test_sim3_like_c1_with_sun_and_subsume(capi):
    # BASELINE config 1 geometry: sun (Subsume, r=500) + two dense elastic clusters
    b = clouds.config("C1", n=601)
    (s,) = run_both(capi, b, 1e-9, 1.0)
    assert np.array_equal(s["pairs"], s["ref_pairs"])
    assert len(s["pairs"]) > 100
    assert_forces(*s["forces"], oracle_sim(b.copy()))
    for f in ("vx", "vy", "vz"):
        assert np.allclose(getattr(s["state"], f), getattr(s["ref"], f), rtol=1e-12, atol=1e-4), f


def test_subsume_events_are_reported_to_host(capi):
    from oracle.oracle import EV_SUBSUME
    rng = np.random.default_rng(3)
    n = 400
    b = clouds.uniform_cube(n, 80.0, 1.0, 1e10, vmax=1.0, seed=5)
    b.radius[:] = rng.uniform(0.5, 6.0, n)
    b.behavior[rng.random(n) < 0.3] = SUBSUME
    b.behavior[rng.random(n) < 0.1] = NONE
    o = oracle_sim(b.copy())
    o.compute()
    ref = sorted((int(e["a"]), int(e["b"]), float(e["dist"])) for e in o.events if e["kind"] == EV_SUBSUME)
    assert len(ref) > 10
    sim = capi.Sim(n)
    sim.upload(b)
    sim.step(1e-6, 1.0)
    hev = sim.host_events()
    got = [(int(e["a"]), int(e["b"]), float(e["dist"])) for e in hev if e["kind"] == capi.EV_SUBSUME]
    assert got == ref
    assert all(int(e["applied"]) == 1 for e in hev)
    assert np.array_equal(sim.pairs(), o.collision_pairs())
    sim.close()


def test_subsume_is_resolved_on_device_in_event_order(capi):
    """ResolveSubsume (body.go:228-244) inside ProcessMods, interleaved with the elastic events of the
    same cycle in the reference's serial order, and before Update: the swallowed body stops existing
    (no Update), the swallower is kicked with its new mass.  Chains (A swallows B, B swallows C) make
    the order observable.  Masses and Exists bit-exact over several cycles."""
    rng = np.random.default_rng(11)
    n = 600
    b = clouds.uniform_cube(n, 60.0, 1.0, 1e12, vmax=200.0, seed=23)
    b.radius[:] = rng.uniform(0.5, 7.0, n)
    b.mass[:] = rng.uniform(1e11, 1e13, n)
    b.behavior[rng.random(n) < 0.35] = SUBSUME
    b.behavior[rng.random(n) < 0.05] = NONE
    o = oracle_sim(b.copy())
    sim = capi.Sim(n)
    sim.upload(b)
    swallowed = 0
    for step in range(4):
        before = o.b.exists.copy()
        o.compute()
        o.process_mods()
        o.update(1e-4, 0.9)
        res = sim.step(1e-4, 0.9)
        g = sim.download()
        assert np.array_equal(g.mass, o.b.mass), f"step {step}: mass"
        assert np.array_equal(g.exists, o.b.exists), f"step {step}: Exists"
        assert res.n_subsumed == int((before & ~o.b.exists).sum())
        assert res.n_dead == int((~o.b.exists).sum())
        swallowed += res.n_subsumed
        for f in ("x", "y", "z", "vx", "vy", "vz"):
            a, r = getattr(g, f), getattr(o.b, f)
            assert np.allclose(a, r, rtol=1e-10, atol=1e-10 * np.max(np.abs(r))), f"step {step}: {f}"
    assert swallowed > 20
    # the dead bodies leave with the next compaction, exactly like BodyCollection.Cycle
    n_new, _ = sim.compact()
    assert n_new == int(o.b.exists.sum())
    sim.close()


@pytest.mark.parametrize("n", [500, 20_000], ids=["small-tiles", "large-tiles"])
def test_deleted_bodies_still_take_part_in_subsume_events(capi, n):
    """Bodies set not to exist at the cycle top (RemoveBodies zeroes the mass, mod-body exists=false keeps it;
    computation-runner.go:176-216, body.go:93-96,308-309) stay in the array until Cycle.  The force sweep skips
    them, the collision sweep does not (body.go:172-186) and ResolveSubsume has no Exists gate (:228-244): a
    deleted body with the larger radius still swallows a live one, a deleted body inside a live subsumer adds
    the mass it kept.  Events, masses and Exists bit-exact against the oracle's canonical stream."""
    from oracle.oracle import EV_SUBSUME, OPT_CANONICAL
    rng = np.random.default_rng(41)
    side = 60.0 * (n / 500) ** (1 / 3)
    b = clouds.uniform_cube(n, side, 1.0, 1e12, vmax=100.0, seed=43)
    b.radius[:] = rng.uniform(0.5, 7.0, n)
    b.mass[:] = rng.uniform(1e11, 1e13, n)
    b.behavior[rng.random(n) < 0.35] = SUBSUME
    b.behavior[rng.random(n) < 0.05] = NONE
    dead = rng.random(n) < 0.15
    b.flags[dead] &= ~np.uint8(F_EXISTS)
    b.mass[dead & (rng.random(n) < 0.5)] = 0.0        # SetNotExists; the others keep their mass (exists=false mod)
    o = oracle_sim(b.copy())
    sim = capi.Sim(n)
    sim.upload(b)
    o.compute(opts=OPT_CANONICAL)
    ref = sorted((int(e["a"]), int(e["b"]), float(e["dist"])) for e in o.events if e["kind"] == EV_SUBSUME)
    with_dead = [e for e in ref if dead[e[0]] or dead[e[1]]]
    assert len(with_dead) > 10 and any(dead[a] for a, _, _ in with_dead) and any(dead[b_] for _, b_, _ in with_dead)
    ref_pairs = o.collision_pairs()
    before = o.b.exists.copy()
    o.process_mods()
    o.update(1e-4, 0.9)
    res = sim.step(1e-4, 0.9)
    hev = sim.host_events()
    assert [(int(e["a"]), int(e["b"]), float(e["dist"])) for e in hev if e["kind"] == capi.EV_SUBSUME] == ref
    assert np.array_equal(sim.pairs(), ref_pairs)     # no collision event names a deleted body
    g = sim.download()
    assert np.array_equal(g.mass, o.b.mass) and np.array_equal(g.exists, o.b.exists)
    assert res.n_subsumed == int((before & ~o.b.exists).sum()) > 0
    for f in ("x", "vx", "vz"):
        a, r = getattr(g, f), getattr(o.b, f)
        assert np.allclose(a, r, rtol=1e-10, atol=1e-10 * np.max(np.abs(r))), f
    sim.close()


def test_subsume_report_only_when_step_is_not_applied(capi):
    b = BodyArrays.from_fields([0, 1, 50], [0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0],
                               [5e10, 1e10, 1e10], [4.0, 1.0, 1.0])
    b.behavior[:] = SUBSUME
    sim = capi.Sim(3)
    sim.upload(b)
    res = sim.step(1e-3, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE)
    hev = sim.host_events()
    assert [(int(e["a"]), int(e["b"]), int(e["applied"])) for e in hev] == [(0, 1, 0), (0, 1, 0)]
    assert res.n_subsumed == 0 and res.n_pairs == 0
    g = sim.download()
    assert np.array_equal(g.mass, b.mass) and g.exists.all()
    res = sim.step(1e-3, 1.0)
    g = sim.download()
    assert res.n_subsumed == 1 and list(g.mass) == [6e10, 0.0, 1e10] and list(g.exists) == [True, False, True]
    assert [int(e["applied"]) for e in sim.host_events()] == [1, 1]
    sim.close()


def test_fragment_decisions_are_handed_to_host(capi):
    from oracle.oracle import EV_FRAGMENT
    b = clouds.uniform_cube(300, 40.0, 2.0, 1e12, vmax=500.0, seed=13)
    b.behavior[::3] = FRAGMENT
    b.frag_factor[:] = 0.05
    b.frag_step[:] = 100.0
    o = oracle_sim(b.copy())
    o.compute()
    o.process_mods()
    ref = sorted((int(e["a"]), int(e["b"])) for e in o.host_events if e["kind"] == EV_FRAGMENT)
    assert len(ref) > 0
    sim = capi.Sim(b.n)
    sim.upload(b)
    sim.step(1e-6, 1.0, capi.STEP_COLLISIONS)
    hev = sim.host_events()
    got = sorted((int(e["a"]), int(e["b"])) for e in hev if e["kind"] == capi.EV_FRAGMENT)
    assert got == ref
    sim.close()


def test_fragmentation_is_recorded_at_event_time(capi):
    """initiateFragmentation (fragcalc.go:66-83) reads the body's Mass and position while the queue is processed:
    after the subsumes handled earlier in the queue, before Update moves the body.  One NB_EV_FRAG_INIT record per
    call, in the reference's handling order, with that mass; nb_get_cycle_top_positions returns the positions
    ProcessMods saw.  Subsume bodies in the scene make the event-time mass differ from the mass after the cycle."""
    from oracle.oracle import EV_FRAG_INIT
    rng = np.random.default_rng(19)
    n = 500
    b = clouds.uniform_cube(n, 40.0, 1.0, 1e12, vmax=500.0, seed=29)
    b.radius[:] = rng.uniform(0.8, 4.0, n)
    b.mass[:] = rng.uniform(1e11, 1e13, n)
    b.behavior[::3] = FRAGMENT
    b.behavior[1::7] = SUBSUME
    b.frag_factor[:] = 0.05
    b.frag_step[:] = 100.0
    x0 = b.x.copy()
    o = oracle_sim(b.copy())
    o.compute()
    o.process_mods()
    ref = [(int(e["a"]), int(e["b"]), int(e["f2"]), float(e["dist"]), float(e["f1"]))
           for e in o.host_events if e["kind"] == EV_FRAG_INIT]
    assert len(ref) > 10
    sim = capi.Sim(n)
    sim.upload(b)
    sim.step(1e-6, 1.0)
    got = [(int(e["a"]), int(e["b"]), int(e["applied"]), float(e["dist"]), float(e["f1"]))
           for e in sim.host_events() if e["kind"] == capi.EV_FRAG_INIT]
    # same calls, same order (the reference's handling order), the mass of that moment bit for bit
    assert [g[:4] for g in got] == [r[:4] for r in ref]
    assert np.allclose([g[4] for g in got], [r[4] for r in ref], rtol=1e-9)
    # some body's event-time mass is not the mass it ends the cycle with (a subsume later in the queue)
    end_mass = sim.download().mass
    assert any(end_mass[a] != m for a, _, _, m, _ in got) or not (b.behavior == SUBSUME).any()
    # the positions ProcessMods saw: the uploaded ones, not the ones Update produced
    px, _, _ = sim.cycle_top_positions(0, n)
    assert np.array_equal(px, x0) and not np.array_equal(sim.download().x, x0)
    sim.close()


def test_dead_and_fragmenting_bodies(capi):
    b = clouds.uniform_cube(500, 50.0, 1.5, 1e12, vmax=5.0, seed=17)
    b.flags[[3, 77, 250]] = 0          # dead: no force from or on them, no events
    b.mass[[3, 77, 250]] = 0
    b.flags[[10, 11]] |= F_FRAGMENTING  # skipped as i and as j
    b.x[77] = np.nan                    # NaN-culled body awaiting compaction
    (s,) = run_both(capi, b, 1e-3, 1.0)
    assert np.array_equal(s["pairs"], s["ref_pairs"])
    fx, fy, fz = s["forces"]
    live = (b.flags & F_EXISTS != 0) & (b.flags & F_FRAGMENTING == 0)
    o = oracle_sim(b.copy())
    ex, ey, ez, fn = o.compute_exact()
    err = np.abs(fx - ex)[live]
    assert np.all(err <= FORCE_TOL * fn[live])
    assert np.all(fx[~live] == 0)
    assert s["res"].n_dead == 3
    xyz, ex_flags = s["render"]
    assert np.array_equal(ex_flags, s["ref_render"][1])
    assert np.allclose(xyz[ex_flags == 1], s["ref_render"][0][ex_flags == 1], rtol=1e-6)
    assert np.all(xyz[ex_flags == 0] == 0)


def test_mixed_radii_threshold_screen(capi):
    # a huge body in one tile must not perturb exactness of the screen for the others
    rng = np.random.default_rng(23)
    n = 1200
    b = clouds.uniform_cube(n, 300.0, 1.0, 1e14, vmax=1.0, seed=29)
    b.radius[:] = rng.uniform(0.1, 8.0, n)
    b.radius[600] = 120.0
    b.mass[600] = 1e20
    (s,) = run_both(capi, b, 1e-5, 1.0)
    assert np.array_equal(s["pairs"], s["ref_pairs"])
    assert_forces(*s["forces"], oracle_sim(b.copy()))


def test_touching_and_coincident_predicate_edges(capi):
    # dist == r1+r2 exactly ⇒ collision and no force; one ulp more ⇒ force and no collision;
    # coincident centres ⇒ collision with NaN velocities ⇒ NaN cull in the same step (KAT-6)
    b = BodyArrays.from_fields([0, 3, 0, 50, 50], [0, 0, 3.0000000000000004, 50, 50], [0, 0, 0, 50, 50],
                               [0.5, -0.5, 0, 1, 3], [0, 0, 0, 2, 2], [0, 0, 0, 3, 1],
                               [1e10] * 5, [1.5, 1.5, 1.5, 2, 2])
    (s,) = run_both(capi, b, 1e-3, 1.0)
    assert np.array_equal(s["pairs"], s["ref_pairs"])
    assert [tuple(p) for p in s["pairs"]] == [(0, 1), (1, 0), (3, 4), (4, 3)]
    assert np.array_equal(s["state"].exists, s["ref"].exists)
    assert list(s["state"].exists) == [True, True, True, False, False]
    assert s["res"].n_culled == 2


def test_multi_step_drift(capi):
    # short trajectory: the documented drift bound is 1e-9 relative to the cloud size over 25 steps
    b = clouds.uniform_cube(800, 120.0, 1.2, 1e13, vmax=20.0, seed=31)
    steps = run_both(capi, b, 1e-2, 1.0, steps=25)
    for k, s in enumerate(steps):
        assert np.array_equal(s["pairs"], s["ref_pairs"]), f"pair set diverged at step {k}"
    last = steps[-1]
    for f in ("x", "y", "z"):
        assert np.max(np.abs(getattr(last["state"], f) - getattr(last["ref"], f))) <= 1e-9 * 120.0
    for f in ("vx", "vy", "vz"):
        assert np.max(np.abs(getattr(last["state"], f) - getattr(last["ref"], f))) <= 1e-9 * 20.0


def test_restitution_applies_from_previous_update(capi):
    # Body.r is set by Update (body.go:129): a new R acts on collisions of the NEXT cycle
    b = BodyArrays.from_fields([0, 1.9], [0, 0], [0, 0], [1, -1], [0.1, 0], [0, 0.2], [2, 1], [1, 1])
    steps = run_both(capi, b, 1e-4, 0.5, steps=3)
    for s in steps:
        for f in ("vx", "vy", "vz"):
            assert np.allclose(getattr(s["state"], f), getattr(s["ref"], f), rtol=1e-12, atol=1e-14)
        assert np.array_equal(s["state"].rest, s["ref"].rest)


# ---------------------------------------------------------------- state sync (cycle top)
def test_empty_collection(capi):
    sim = capi.Sim(16)
    sim.upload(BodyArrays(0))
    res = sim.step(1e-3, 1.0)
    assert res.n_bodies == 0 and res.n_pairs == 0
    assert sim.pairs().shape == (0, 2)
    sim.close()


def test_patch_append_compact_semantics(capi):
    # TestRemove / TestAdds / TestMod semantics (cmd/body/body_collection_test.go:73-88,268-281)
    n = 1000
    b = clouds.uniform_cube(n, 400.0, 1.0, 1e12, seed=37)
    sim = capi.Sim(n + 10)
    sim.upload(b)
    # SetNotExists on three bodies (mass 0, exists false)
    for i in (5, 500, 999):
        sim.patch(i, 1, mass=np.zeros(1), flags=np.zeros(1, dtype=np.uint8))
    add = BodyArrays.from_fields([7.0], [8.0], [9.0], [1.0], [2.0], [3.0], [5e11], [2.0])
    sim.append(add, R=0.75)
    assert sim.count() == n + 1
    new_n, old = sim.compact()
    keep = [i for i in range(n) if i not in (5, 500, 999)] + [n]
    assert new_n == n - 2 and list(old) == keep
    got = sim.download()
    assert np.array_equal(got.x[:-1], b.x[keep[:-1]]) and got.x[-1] == 7.0
    assert got.rest[-1] == 0.75 and np.all(got.rest[:-1] == 1.0)
    # ApplyMods: x= and collision=subsume on one body
    sim.patch(10, 1, x=np.array([41.0]), behavior=np.array([SUBSUME], dtype=np.uint8))
    got = sim.download()
    assert got.x[10] == 41.0 and got.behavior[10] == SUBSUME
    # the stepped state after compaction still matches an oracle fed the same bodies
    o = oracle_sim(got.copy())
    o.step(1e-3, 1.0)
    sim.step(1e-3, 1.0)
    g2 = sim.download()
    assert np.allclose(g2.x, o.b.x, rtol=1e-12, atol=1e-9)
    sim.close()


def test_pair_overflow_leaves_state_untouched(capi):
    b = clouds.uniform_cube(600, 30.0, 2.0, 1e12, vmax=10.0, seed=41)
    sim = capi.Sim(b.n, pair_capacity=16)
    sim.upload(b)
    with pytest.raises(capi.NbError) as ei:
        sim.step(1e-3, 1.0)
    assert ei.value.code == capi.NB_ERR_PAIR_OVERFLOW
    got = sim.download()
    assert np.array_equal(got.x, b.x) and np.array_equal(got.vx, b.vx)
    sim.close()


def test_capacity_and_argument_errors(capi):
    sim = capi.Sim(4)
    with pytest.raises(capi.NbError) as ei:
        sim.upload(BodyArrays(5))
    assert ei.value.code == capi.NB_ERR_CAPACITY
    sim.upload(BodyArrays(4))
    with pytest.raises(capi.NbError):
        sim.append(BodyArrays(1), 1.0)
    with pytest.raises(capi.NbError):
        sim.patch(3, 2, x=np.zeros(2))
    sim.close()


# ---------------------------------------------------------------- size-independent properties at scale
def test_register_blocking_does_not_change_bits(capi):
    # R (i-bodies per thread) is a launch-shape choice; per-body summation order is a function of n only
    b = clouds.uniform_cube(20_000, 900.0, 1.0, 1e14, vmax=5.0, seed=43)
    outs = []
    for R in ("1", "2", "4"):
        os.environ["NB_FORCE_R"] = R
        sim = capi.Sim(b.n)
        sim.upload(b)
        sim.step(1e-3, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE)
        outs.append((sim.forces(), sim.pairs()))
        sim.close()
    os.environ.pop("NB_FORCE_R")
    for (f, p) in outs[1:]:
        assert all(np.array_equal(a.view(np.uint64), r.view(np.uint64)) for a, r in zip(f, outs[0][0]))
        assert np.array_equal(p, outs[0][1])


def test_c3_scale_properties_and_sampled_parity(capi):
    # BASELINE config 3 at full size: 100,000-body cube with elastic collisions
    b = clouds.config("C3")
    sim = capi.Sim(b.n)
    sim.upload(b)
    res = sim.step(1e-9, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE)
    fx, fy, fz = sim.forces()
    pairs = sim.pairs()
    sim.close()
    # pair set symmetric and free of self pairs
    assert len(pairs) > 0 and len(pairs) % 2 == 0
    assert set(map(tuple, pairs)) == set((j, i) for i, j in pairs)
    assert not np.any(pairs[:, 0] == pairs[:, 1])
    # Newton's third law, normwise
    fmag = np.sqrt(fx * fx + fy * fy + fz * fz).sum()
    assert abs(fx.sum()) + abs(fy.sum()) + abs(fz.sum()) <= 1e-10 * fmag
    # sampled rows against the oracle: 48 random bodies + every body that collides
    rng = np.random.default_rng(0)
    rows = np.unique(np.concatenate([rng.integers(0, b.n, 24), pairs[:16, 0]]))
    o = oracle_sim(b.copy())
    ref_pairs = []
    for i in rows:
        o.compute(int(i), int(i) + 1)
        ref_pairs += [tuple(p) for p in o.collision_pairs()]
        ex, ey, ez, fn = o.compute_exact(int(i), int(i) + 1)
        err = max(abs(fx[i] - ex[i]), abs(fy[i] - ey[i]), abs(fz[i] - ez[i]))
        assert err <= FORCE_TOL * fn[i], (i, err / fn[i])
    rowset = set(rows.tolist())
    got_rows = [tuple(p) for p in pairs if p[0] in rowset]
    assert got_rows == sorted(ref_pairs)
    assert res.n_pairs == len(pairs)


def test_c4_full_size_sampled_parity(capi):
    # BASELINE config 4: 1,000,000-body uniform sphere, elastic; one full step on one GPU,
    # checked on sampled rows (the oracle cannot do 1e12 pairs) and through global properties
    b = clouds.config("C4")
    sim = capi.Sim(b.n)
    sim.upload(b)
    res = sim.step(1e-9, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE)
    fx, fy, fz = sim.forces()
    pairs = sim.pairs()
    sim.close()
    assert len(pairs) > 0 and set(map(tuple, pairs)) == set((j, i) for i, j in pairs)
    fmag = np.sqrt(fx * fx + fy * fy + fz * fz).sum()
    assert abs(fx.sum()) + abs(fy.sum()) + abs(fz.sum()) <= 1e-9 * fmag
    rng = np.random.default_rng(1)
    rows = np.unique(np.concatenate([rng.integers(0, b.n, 3), pairs[:3, 0]]))
    o = oracle_sim(b.copy())
    ref_pairs = []
    for i in rows:
        o.compute(int(i), int(i) + 1)
        ref_pairs += [tuple(p) for p in o.collision_pairs()]
        ex, ey, ez, fn = o.compute_exact(int(i), int(i) + 1)
        err = max(abs(fx[i] - ex[i]), abs(fy[i] - ey[i]), abs(fz[i] - ez[i]))
        assert err <= FORCE_TOL * fn[i], (i, err / fn[i])
    rowset = set(rows.tolist())
    assert [tuple(p) for p in pairs if p[0] in rowset] == sorted(ref_pairs)
    assert res.n_pairs == len(pairs)


def test_c2_full_size_collisions_off(capi):
    # BASELINE config 2 at full size: 10,000-body uniform sphere, behaviour None (no events), 3 steps
    b = clouds.config("C2")
    steps = run_both(capi, b, 1e-9, 1.0, opts=0, steps=3)
    o = oracle_sim(b.copy())
    fx, fy, fz = steps[0]["forces"]
    for i0 in (0, 5000, 9900):  # __float128 adjudicator on 300 sampled rows
        ex, ey, ez, fn = o.compute_exact(i0, i0 + 100)
        sl = slice(i0, i0 + 100)
        err = np.max(np.abs(np.stack([fx[sl] - ex[sl], fy[sl] - ey[sl], fz[sl] - ez[sl]])), axis=0)
        assert np.all(err <= FORCE_TOL * fn[sl])
    last = steps[-1]
    assert last["res"].n_pairs == 0 and len(last["hev"]) == 0
    for f in ("x", "y", "z", "vx", "vy", "vz"):
        r = getattr(last["ref"], f)
        assert np.allclose(getattr(last["state"], f), r, rtol=1e-10, atol=1e-12 * np.max(np.abs(r))), f


def test_c3_dense_full_size_full_cycle(capi):
    # BASELINE config 3, collision-dense variant (r x 4): one full cycle of 100,000 bodies against the
    # threaded oracle: the complete pair list bit-exact, every post-step velocity within tolerance
    b = clouds.config("C3dense")
    o = oracle_sim(b.copy())
    o.compute(workers=os.cpu_count() or 8)
    ref_pairs = o.collision_pairs()
    assert len(ref_pairs) > 5000
    o.process_mods()
    o.update(1e-9, 1.0)
    sim = capi.Sim(b.n)
    sim.upload(b)
    res = sim.step(1e-9, 1.0)
    assert np.array_equal(sim.pairs(), ref_pairs)
    got = sim.download()
    assert res.n_pairs == len(ref_pairs) and res.resolve_rounds >= 2
    for f in ("vx", "vy", "vz"):
        assert np.max(np.abs(getattr(got, f) - getattr(o.b, f))) <= 1e-11 * 1e8, f
    for f in ("x", "y", "z"):
        assert np.max(np.abs(getattr(got, f) - getattr(o.b, f))) <= 1e-11 * 2000.0, f
    sim.close()


def test_async_step_and_sync(capi):
    # NB_STEP_ASYNC enqueues the cycle; nb_sync collects the same result a synchronous step gives
    b = clouds.uniform_cube(2000, 150.0, 1.5, 1e13, vmax=20.0, seed=51)
    s1, s2 = capi.Sim(b.n), capi.Sim(b.n)
    s1.upload(b)
    s2.upload(b)
    r1 = s1.step(1e-3, 1.0)
    s2.step(1e-3, 1.0, capi.STEP_DEFAULT | capi.STEP_ASYNC)
    r2 = s2.sync()
    assert (r1.n_pairs, r1.n_resolved, r1.n_dead) == (r2.n_pairs, r2.n_resolved, r2.n_dead) and r1.n_pairs > 0
    a, c = s1.download(), s2.download()
    assert np.array_equal(a.x.view(np.uint64), c.x.view(np.uint64))
    assert np.array_equal(a.vx.view(np.uint64), c.vx.view(np.uint64))
    s1.close()
    s2.close()


def test_no_resolve_hands_pairs_to_host(capi):
    # NB_STEP_NO_RESOLVE: detection only — velocities are those of a collision-free cycle, the pair
    # list is still complete, collided flags stay clear
    b = clouds.uniform_cube(1500, 60.0, 2.5, 1e12, vmax=100.0, seed=21)
    o = oracle_sim(b.copy())
    o.compute()
    ref_pairs = o.collision_pairs()
    o.update(1e-4, 1.0)            # no ProcessMods
    sim = capi.Sim(b.n)
    sim.upload(b)
    res = sim.step(1e-4, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_RESOLVE)
    assert res.n_resolved == 0 and res.n_pairs == len(ref_pairs)
    assert np.array_equal(sim.pairs(), ref_pairs)
    g = sim.download()
    for f in ("vx", "x"):
        r = getattr(o.b, f)
        assert np.allclose(getattr(g, f), r, rtol=1e-11, atol=1e-12 * np.max(np.abs(r)))
    sim.close()


def test_no_integrate_is_side_effect_free_on_state(capi):
    b = clouds.uniform_cube(800, 60.0, 2.0, 1e12, vmax=10.0, seed=5)
    sim = capi.Sim(b.n)
    sim.upload(b)
    sim.step(1e-3, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE)
    g = sim.download()
    for f in ("x", "y", "z", "vx", "vy", "vz", "rest"):
        assert np.array_equal(getattr(g, f), getattr(b, f)), f
    assert np.array_equal(g.flags, b.flags)
    assert len(sim.pairs()) > 0
    sim.close()


def test_range_calls_and_force_roundtrip(capi):
    """The round-2 ABI additions on one GPU: nb_upload_shard (== nb_upload on a single handle),
    nb_download_state_range / nb_download_render_range, nb_set_forces, nb_comm_mode, and their argument checks."""
    b = clouds.uniform_cube(700, 60.0, 2.0, 1e15, vmax=10.0, seed=31)
    ref = capi.Sim(b.n)
    ref.upload(b)
    sim = capi.Sim(b.n)
    sim.upload_shard(b.n, 0, b.n, b.x, b.y, b.z, b.vx, b.vy, b.vz, b.mass, b.radius, rest=b.rest,
                     ff=b.frag_factor, fs=b.frag_step, behavior=b.behavior, flags=b.flags)
    assert sim.comm_mode() == capi.COMM_SINGLE
    with pytest.raises(capi.NbError):      # not this handle's i-range
        sim.upload_shard(b.n, 10, b.n - 10, *[a[10:] for a in (b.x, b.y, b.z, b.vx, b.vy, b.vz, b.mass, b.radius)])
    for s in (ref, sim):
        s.step(1e-3, 1.0)
    full = ref.download()
    x, vz = np.zeros(100), np.zeros(100)
    sim.download_range_into(250, 100, x=x, vz=vz)
    assert np.array_equal(x, full.x[250:350]) and np.array_equal(vz, full.vz[250:350])
    xyz, ex = np.zeros((100, 3), dtype=np.float32), np.zeros(100, dtype=np.uint8)
    sim.render_range(250, 100, xyz, ex)
    rxyz, rex = ref.render()
    assert np.array_equal(xyz, rxyz[250:350]) and np.array_equal(ex, rex[250:350])
    with pytest.raises(capi.NbError):
        sim.download_range_into(650, 100, x=x)
    # forces survive a host round trip (what a re-upload of a running collection does for fragmenting bodies)
    fx, fy, fz = ref.forces()
    sim.set_forces(0, b.n, np.zeros(b.n), np.zeros(b.n), np.zeros(b.n))
    assert not sim.forces()[0].any()
    sim.set_forces(5, 20, fx[5:25], fy[5:25], fz[5:25])
    g = sim.forces()
    assert np.array_equal(g[0][5:25], fx[5:25]) and np.array_equal(g[2][5:25], fz[5:25]) and not g[0][25:].any()
    with pytest.raises(capi.NbError):
        sim.set_forces(690, 20, fx[:20], fy[:20], fz[:20])
    with pytest.raises(capi.NbError):      # compaction invalidates the cycle-top positions
        sim.compact()
        sim.cycle_top_positions(0, 10)
    ref.close()
    sim.close()


def test_resolve_cluster_equals_single_cta(capi, monkeypatch):
    """K3 runs as a thread-block cluster of 8 CTAs when the event list is longer than one CTA (round 2); which
    events share a round — and therefore every bit of the result — must not depend on the number of CTAs.
    Dense cloud: several thousand events in a dozen rounds, subsume chains and fragment decisions included."""
    rng = np.random.default_rng(37)
    n = 6000
    b = clouds.uniform_cube(n, 95.0, 2.5, 1e12, vmax=100.0, seed=39)
    b.radius[:] = rng.uniform(1.0, 4.0, n)
    b.behavior[rng.random(n) < 0.1] = SUBSUME
    b.behavior[rng.random(n) < 0.1] = FRAGMENT
    b.frag_factor[:] = 0.05
    b.frag_step[:] = 100.0
    outs = []
    for cluster in ("1", "8", "3"):
        monkeypatch.setenv("NB_RES_CLUSTER", cluster)      # read at nb_create
        sim = capi.Sim(n)
        sim.upload(b)
        log = []
        for _ in range(3):
            res = sim.step(1e-4, 0.9)
            st = sim.download()
            log.append((res.n_pairs, res.n_resolved, res.n_subsumed, res.resolve_rounds, st.vx.copy(), st.vz.copy(),
                        st.mass.copy(), st.flags.copy(), st.behavior.copy(),
                        [tuple(e) for e in sim.host_events().tolist()]))
        sim.close()
        outs.append(log)
    assert outs[0][0][0] > 2000 and outs[0][0][3] > 3       # longer than one CTA, several rounds
    for other in outs[1:]:
        for a, c in zip(outs[0], other):
            assert a[:4] == c[:4]
            for x, y in zip(a[4:9], c[4:9]):
                assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
            assert a[9] == c[9]


def test_short_event_lists_use_the_shared_memory_schedule_with_the_same_bits(capi, monkeypatch):
    """Event lists of up to 64 entries are scheduled in shared memory by the first CTA of K3 (no per-body lists in
    global memory); the readiness rule is the same, so rounds, counters, host records and every state bit must equal
    the general path's (NB_RES_FAST=0).  Small dense clouds: chains through shared bodies (several rounds), subsume
    and fragment decisions, lists just under, at and just over the limit."""
    rng = np.random.default_rng(53)
    seen_rounds, seen_sizes = set(), []
    for n, side, seed in ((150, 30.0, 1), (260, 38.0, 2), (400, 48.0, 3), (90, 20.0, 4), (320, 40.0, 5)):
        b = clouds.uniform_cube(n, side, 1.2, 1e12, vmax=200.0, seed=60 + seed)
        b.radius[:] = rng.uniform(0.6, 2.2, n)
        b.behavior[rng.random(n) < 0.12] = SUBSUME
        b.behavior[rng.random(n) < 0.12] = FRAGMENT
        b.frag_factor[:] = 0.05
        b.frag_step[:] = 100.0
        outs = []
        for fast in ("1", "0"):
            monkeypatch.setenv("NB_RES_FAST", fast)          # read at nb_create
            sim = capi.Sim(n)
            sim.upload(b)
            log = []
            for _ in range(4):
                res = sim.step(2e-4, 0.85)
                st = sim.download()
                log.append((res.n_pairs, res.n_resolved, res.n_subsumed, res.resolve_rounds, res.n_host_events,
                            st.vx.copy(), st.vy.copy(), st.vz.copy(), st.mass.copy(), st.flags.copy(),
                            st.behavior.copy(), [tuple(e) for e in sim.host_events().tolist()],
                            sorted(map(tuple, sim.pairs().tolist()))))
            sim.close()
            outs.append(log)
        for a, c in zip(*outs):
            assert a[:5] == c[:5]
            for x, y in zip(a[5:11], c[5:11]):
                assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
            assert a[11] == c[11] and a[12] == c[12]
            seen_rounds.add(a[3])
            seen_sizes.append(a[0] + a[2])
    assert any(0 < k <= 64 for k in seen_sizes) and any(k > 64 for k in seen_sizes)   # both sides of the limit
    assert max(seen_rounds) >= 3                                                       # chains, not just isolated pairs
