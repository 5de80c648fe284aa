"""GPU parity of the elastic collision response across its whole conditioning range.

calcElasticCollision (cmd/body/collisioncalc.go:42-186) derives the approach angle as
thetav = Acos(vz1r / v) and the impact angle as alpha = Asin(-dr): the last bit of their
arguments is amplified by 1/thetav (nearly head-on approaches) and 1/sqrt(1 - dr^2) (grazing
ones).  tests/test_gomath.py measures that for two CPU libms; this file holds the device to the
same conditioning-aware bound — against the oracle on glibc AND on the restated Go library —
and to the fixed 1e-13 wherever the formula is well conditioned.
"""
import math

import numpy as np
import pytest

from nbodygo_b200.bodies import BodyArrays

pytestmark = pytest.mark.gpu

EPS = 2.0 ** -52
SPEED = 1.0e6
ANGLES = (0.0, 1e-9, 3e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 0.1, 0.5, 1.0)


def isolated_pairs(per_angle=24, seed=17):
    """Pairs of overlapping unit spheres, 1e4 apart from each other (no cross-pair overlap),
    body 2k approaching body 2k+1 at angle ANGLES[...] to the line of centres."""
    rng = np.random.default_rng(seed)
    P, V, M, th = [], [], [], []
    k = 0
    for ang in ANGLES:
        for _ in range(per_angle):
            u = rng.normal(size=3)
            u /= np.linalg.norm(u)
            w = np.cross(u, rng.normal(size=3))
            w /= np.linalg.norm(w)
            centre = np.array([(k % 16) * 1e4, ((k // 16) % 16) * 1e4, (k // 256) * 1e4])
            pa = centre + rng.uniform(-10, 10, 3)
            pb = pa + rng.uniform(0.3, 1.9) * u
            vb = rng.normal(0, 1e3, 3)
            va = vb + SPEED * (math.cos(ang) * u + math.sin(ang) * w)
            P += [pa, pb]
            V += [va, vb]
            M += [10 ** rng.uniform(10, 12), 10 ** rng.uniform(10, 12)]
            th.append(ang)
            k += 1
    P, V = np.array(P), np.array(V)
    n = len(P)
    b = BodyArrays.from_fields(P[:, 0], P[:, 1], P[:, 2], V[:, 0], V[:, 1], V[:, 2], np.array(M), np.ones(n))
    return b, np.array(th)


def conditioning(b):
    """1/thetav + 1/sqrt(1 - dr^2) per pair (2k, 2k+1), from the inputs in double."""
    c = []
    for k in range(0, b.n, 2):
        vrel = np.array([b.vx[k] - b.vx[k + 1], b.vy[k] - b.vy[k + 1], b.vz[k] - b.vz[k + 1]])
        axis = np.array([b.x[k + 1] - b.x[k], b.y[k + 1] - b.y[k], b.z[k + 1] - b.z[k]])
        d, s = np.linalg.norm(axis), np.linalg.norm(vrel)
        th = math.acos(max(-1.0, min(1.0, float(vrel @ axis) / s / d)))
        dr = d * math.sin(th) / (b.radius[k] + b.radius[k + 1])
        # the computed thetav is quantised near 0: Acos(1 - 2^-53) = 1.5e-8 is its smallest non-zero value
        c.append(1 / max(th, 1.49e-8) + 1 / math.sqrt(max(1 - dr * dr, 1e-16)))
    return np.array(c)


def oracle_velocities(b, backend):
    from oracle import oracle
    from oracle.oracle import OracleSim
    prev = oracle.set_math(backend)
    try:
        o = OracleSim(b.copy())
        o.compute()
        pairs = o.collision_pairs()
        o.process_mods()
        o.update(1e-12, 1.0)
    finally:
        oracle.set_math(prev)
    return pairs, np.stack([o.b.vx, o.b.vy, o.b.vz], axis=1)


@pytest.mark.parametrize("backend", ["libm", "go"])
def test_collision_response_within_eps_times_condition_number(backend):
    from nbodygo_b200 import capi
    from oracle.oracle import MATH_GO, MATH_LIBM
    b, th = isolated_pairs()
    cond = conditioning(b)
    ref_pairs, ref_v = oracle_velocities(b, MATH_GO if backend == "go" else MATH_LIBM)
    sim = capi.Sim(b.n)
    sim.upload(b)
    res = sim.step(1e-12, 1.0)
    got = sim.download()
    pairs = sim.pairs()
    sim.close()
    # every pair overlaps and nothing else does: (2k,2k+1) and (2k+1,2k), bit-exact set
    assert np.array_equal(pairs, ref_pairs) and len(pairs) == b.n
    assert res.n_resolved >= b.n // 2
    v = np.stack([got.vx, got.vy, got.vz], axis=1)
    dv = np.abs(v - ref_v).max(axis=1).reshape(-1, 2).max(axis=1) / SPEED
    ratio = dv / (EPS * cond)
    assert ratio.max() <= 8, (float(ratio.max()), float(th[ratio.argmax()]))
    generic = th >= 0.1
    assert dv[generic].max() <= 1e-13, float(dv[generic].max())
    # the tiny deflection angle is the only ill-conditioned quantity: momentum is conserved throughout
    v0 = np.stack([b.vx, b.vy, b.vz], axis=1)
    p0 = (b.mass[:, None] * v0).reshape(-1, 2, 3).sum(axis=1)
    p1 = (b.mass[:, None] * v).reshape(-1, 2, 3).sum(axis=1)
    mscale = b.mass.reshape(-1, 2).max(axis=1)[:, None] * SPEED
    assert np.max(np.abs(p1 - p0) / mscale) <= 1e-14
    ke0 = (0.5 * b.mass * (v0 * v0).sum(axis=1)).reshape(-1, 2).sum(axis=1)
    ke1 = (0.5 * b.mass * (v * v).sum(axis=1)).reshape(-1, 2).sum(axis=1)
    assert np.max(np.abs(ke1 - ke0) / ke0) <= 1e-13
