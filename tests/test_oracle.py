"""CPU tests: the C oracle against the KATs of SURVEY.md §8c and the independent
pure-Python restatement in tests/golden/ (bit-exact)."""
import numpy as np
import pytest

from helpers import bits, load_golden, same_bits, scene_bodies, scene_step_arrays, unhex
from nbodygo_b200.bodies import BodyArrays, ELASTIC, F_EXISTS, F_COLLIDED, NONE, SUBSUME
from oracle.oracle import EV_COLLISION, EV_SUBSUME, OPT_DEAD_J, OPT_SELF_PAIRS, OracleSim

KIND = {"collision": EV_COLLISION, "subsume": EV_SUBSUME}


@pytest.mark.parametrize("scene", load_golden(), ids=lambda s: s["name"])
def test_oracle_matches_python_restatement(scene):
    b = scene_bodies(scene)
    o = OracleSim(b)
    ts, R = unhex(scene["ts"]), unhex(scene["R"])
    for step in scene["steps"]:
        exp = scene_step_arrays(step)
        o.compute()
        live = b.exists
        got_f = np.stack([o.fx, o.fy, o.fz], axis=1)
        assert same_bits(got_f[live], exp["forces"][live])
        ev = o.events
        assert [(int(e["kind"]), int(e["a"]), int(e["b"])) for e in ev] == \
               [(KIND[k], a, b_) for k, a, b_, _ in step["events"]]
        assert same_bits(ev["dist"], [unhex(d) for *_, d in step["events"]])
        o.process_mods()
        o.update(ts, R)
        for f in ("x", "y", "z", "vx", "vy", "vz", "mass"):
            assert same_bits(getattr(b, f), exp[f]), f
        assert np.array_equal(b.exists, exp["exists"])


def test_kat1_bit_patterns():
    # SURVEY §8c KAT-1 (cmd/runner/workpool_test.go:41-56), hand-derived bit patterns
    b = BodyArrays.from_fields([1, 22], [1, 22], [1, 22], [0, 0], [0, 0], [0, 0], [1, 1], [1, 1])
    o = OracleSim(b)
    o.step(1.0, 1.0)
    assert bits(o.fx)[0] == 0x3D2064B9569010CA
    assert o.fx[1] == -o.fx[0] and o.fy[0] == o.fx[0] and o.fz[0] == o.fx[0]
    assert b.vx[0] == 2.912062242103079e-14
    assert bits(b.x)[0] == 0x3FF0000000000083 and bits(b.x)[1] == 0x4035FFFFFFFFFFF8
    assert b.vx[0] != 0 and b.vx[1] != 0  # the reference test's own assertion


def test_kat2_simtest_bit_patterns():
    # SURVEY §8c KAT-2 (cmd/sim/simgen.go:387-404)
    b = BodyArrays.from_fields(
        [20000, 0, -350, 350], [20000, 0, 350, 350], [20000, 0, 0, 0],
        [-3, 0, 5.3e8, -5.3e8], [-3, 0, -5e8, -5e8], [-5, 0, 0, 0],
        [1, 9e29, 9e29, 9e29], [500, 60, 60, 60])
    b.behavior[0] = SUBSUME
    o = OracleSim(b)
    o.compute()
    exp = [(0xC2346ABB4F10D30D, 0xC2342E7878D7823B, 0xC2346BD150747990),
           (0x0000000000000000, 0x492BFB2D249CA1D7, 0x421AE91707F51239),
           (0x4927E22748A7130A, 0xC91BFB2D249CA1D7, 0x421AE6FB1071CB2E),
           (0xC927E22748A7130A, 0xC91BFB2D249CA1D7, 0x421BDF33296B08D6)]
    for i, (ex, ey, ez) in enumerate(exp):
        assert (bits(o.fx)[i], bits(o.fy)[i], bits(o.fz)[i]) == (ex, ey, ez)
    assert len(o.events) == 0
    o.process_mods()
    o.update(1e-9, 1.0)
    assert (b.vx[2], b.vy[2], b.vz[2]) == (530295898.8243172, -500173333.5181948, 3.2095656422720467e-29)
    assert (b.x[2], b.y[2], b.z[2]) == (-349.4697041011757, 349.4998266664818, 3.209565642272047e-38)


def _pair(p1, v1, m1, r1, p2, v2, m2, r2):
    return BodyArrays.from_fields(*[[a, b_] for a, b_ in zip(p1 + v1, p2 + v2)], [m1, m2], [r1, r2])


def test_kat3_headon():
    o = OracleSim(_pair([0, 0, 0], [1, 0, 0], 1, 1, [1.5, 0, 0], [-1, 0, 0], 1, 1))
    hit, v1, v2, vcm = o.calc_elastic(0, 1)
    assert hit
    assert v1[0] == -1 and v2[0] == 1 and v1[1] == 0 and v2[1] == 0
    assert abs(v1[2]) <= 2e-16 and abs(v2[2]) <= 2e-16  # cos(pi/2) residue, libm-dependent
    assert np.all(vcm == 0)


def test_kat4_oblique():
    o = OracleSim(_pair([0, 0, 0], [3, 2, 1], 2, 1, [1.2, 1.1, 0.9], [-1, 0.5, -2], 3, 1.5))
    hit, v1, v2, vcm = o.calc_elastic(0, 1)
    assert hit
    np.testing.assert_allclose(v1, [-1.1549999476979202, -1.1355002179253342, -2.1162499607734397], rtol=1e-14)
    np.testing.assert_allclose(v2, [1.7699999651319471, 2.5903334786168895, 0.07749997384896012], rtol=1e-13)
    np.testing.assert_allclose(vcm, [0.6, 1.1, -0.8], rtol=1e-15)
    # momentum and kinetic energy conserved
    np.testing.assert_allclose(2 * v1 + 3 * v2, [3, 5.5, -4], rtol=1e-14)
    assert abs(0.5 * 2 * v1 @ v1 + 0.5 * 3 * v2 @ v2 - 21.875) < 1e-13


def test_kat5_mirrored_event_is_noop():
    o = OracleSim(_pair([0, 0, 0], [-1, 0, 0], 1, 1, [1.5, 0, 0], [1, 0, 0], 1, 1))
    hit, *_ = o.calc_elastic(0, 1)
    assert not hit
    o.compute()
    assert len(o.events) == 2
    o.process_mods()
    assert not (o.b.flags & F_COLLIDED).any()
    assert list(o.b.vx) == [-1, 1]


def test_kat6_coincident_nan_cull():
    # cmd/body/body_collection_test.go:319-344: same point ⇒ collided (with NaN velocities),
    # then Update culls both (cmd/body/body.go:134-137)
    o = OracleSim(_pair([500, 500, 500], [1, 2, 3], 5, 2, [500, 500, 500], [3, 2, 1], 7, 2))
    o.compute()
    o.process_mods()
    assert (o.b.flags & F_COLLIDED).all()
    assert np.isnan(o.b.vx).all()
    o.update(1e-3, 1.0)
    assert not o.b.exists.any()
    keep = o.cycle_compact()
    assert o.b.n == 0 and len(keep) == 0


def test_reference_event_stream_options():
    # F5: the raw reference stream holds (i,i) for every elastic body; A2: dead j are not filtered
    b = BodyArrays.from_fields([0, 1, 50], [0, 0, 0], [0, 0, 0], [0] * 3, [0] * 3, [0] * 3, [1] * 3, [1] * 3)
    b.flags[1] = 0
    o = OracleSim(b)
    o.compute()
    assert len(o.events) == 0
    o.compute(opts=OPT_SELF_PAIRS)
    assert [(e["a"], e["b"]) for e in o.events] == [(0, 0), (2, 2)]
    o.compute(opts=OPT_SELF_PAIRS | OPT_DEAD_J)
    assert [(e["a"], e["b"]) for e in o.events] == [(0, 0), (0, 1), (2, 2)]
    # self-pair resolves to a no-op (v == 0 exit, cmd/body/collisioncalc.go:89-92)
    o.process_mods()
    assert not (b.flags & F_COLLIDED).any()


def test_overlapping_bodies_exert_no_gravity():
    # F6 (cmd/body/body.go:219)
    b = BodyArrays.from_fields([0, 1.5], [0, 0], [0, 0], [0, 0], [0, 0], [0, 0], [1e10, 1e10], [1, 1])
    b.behavior[:] = NONE
    o = OracleSim(b)
    o.compute()
    assert o.fx[0] == 0 and o.fx[1] == 0 and len(o.events) == 0


def test_pool_matches_single_worker():
    rng = np.random.default_rng(5)
    n = 700
    b = BodyArrays.from_fields(*(rng.uniform(-30, 30, n) for _ in range(3)),
                               *(rng.uniform(-1, 1, n) for _ in range(3)),
                               rng.uniform(1e9, 1e10, n), rng.uniform(0.5, 2.0, n))
    o1, o2 = OracleSim(b.copy()), OracleSim(b.copy())
    o1.compute()
    o2.compute(workers=7)  # 7 slices of 100
    assert same_bits(o1.fx, o2.fx) and same_bits(o1.fy, o2.fy) and same_bits(o1.fz, o2.fz)
    assert len(o1.events) > 0 and np.array_equal(o1.events, o2.events)
    assert o2.time_slice(0, n, 3) == len(o1.events)


def test_exact_adjudicator_brackets_double_sum():
    rng = np.random.default_rng(9)
    n = 300
    b = BodyArrays.from_fields(*(rng.uniform(-100, 100, n) for _ in range(3)),
                               *(np.zeros(n) for _ in range(3)),
                               rng.uniform(1e20, 1e21, n), np.full(n, 0.5))
    o = OracleSim(b)
    o.compute()
    ex, ey, ez, fn = o.compute_exact()
    err = np.max(np.abs(np.stack([o.fx - ex, o.fy - ey, o.fz - ez])), axis=0)
    assert np.all(err <= 1e-14 * fn)
    # Newton's third law, normwise
    assert abs(ex.sum()) <= 1e-12 * fn.sum()


def test_cycle_compaction_is_stable():
    # TestRemove semantics (cmd/body/body_collection_test.go:73-88)
    b = BodyArrays(10)
    b.x[:] = np.arange(10)
    b.flags[[2, 5, 9]] = 0
    o = OracleSim(b)
    keep = o.cycle_compact()
    assert list(keep) == [0, 1, 3, 4, 6, 7, 8] and list(b.x) == [0, 1, 3, 4, 6, 7, 8]
    assert list(b.id) == [0, 1, 3, 4, 6, 7, 8]
