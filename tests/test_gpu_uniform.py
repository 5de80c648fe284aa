"""The uniform-mass pass of K1 (nb_force.cu): j-chunks whose tiles each hold bodies of one mass run
an instantiation that takes the mass out of the pair loop.  Which chunks those are is a property
of the bodies only; the result must agree with the per-body-mass pass to rounding, with the exact
adjudicator to the stated 1e-12, and pair sets must not depend on it at all."""
import os

import numpy as np
import pytest

from nbodygo_b200 import clouds
from nbodygo_b200.bodies import F_EXISTS

pytestmark = pytest.mark.gpu

N = 40_000   # >= 16384: 256-body tiles


def chunk_start(capi, c):
    """First body of j-chunk c (nb_plan: chunking is a function of n only)."""
    _, _, n_chunks, tiles_per_chunk = capi.plan(N)
    assert c < n_chunks - 1
    return c * tiles_per_chunk * 256


def one_step(capi, b, uniform):
    os.environ["NB_UNIFORM_TILES"] = "1" if uniform else "0"
    try:
        sim = capi.Sim(b.n)
    finally:
        os.environ.pop("NB_UNIFORM_TILES")
    sim.upload(b)
    res = sim.step(1e-9, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE)
    f = np.stack(sim.forces(), axis=1)
    pairs = sim.pairs()
    launches = sim.launch_count()
    sim.close()
    return f, pairs, launches, res


@pytest.fixture(scope="module")
def capi():
    from nbodygo_b200 import capi as c
    c.load()
    return c


def exact_rows(b, rows):
    from oracle.oracle import OracleSim
    o = OracleSim(b.copy())
    fx, fy, fz, fn = o.compute_exact(int(rows[0]), int(rows[-1]) + 1)
    return np.stack([fx, fy, fz], axis=1)[rows], fn[rows]


def test_uniform_pass_agrees_with_general_pass_and_exact_sum(capi):
    b = clouds.uniform_cube(N, 1500.0, 1.2, 1e24, vmax=1e8, seed=5)   # one mass for all
    f0, p0, l0, _ = one_step(capi, b, False)
    f1, p1, l1, _ = one_step(capi, b, True)
    assert np.array_equal(p0, p1) and len(p0) > 0
    assert l1 == l0 + 1          # the second instantiation of K1 over the same grid
    assert not np.array_equal(f0.view(np.uint64), f1.view(np.uint64))   # the pass really ran
    rows = np.arange(1000, 1064)
    ex, fn = exact_rows(b, rows)
    for f in (f0, f1):
        assert np.max(np.abs(f[rows] - ex).max(axis=1) / fn) <= 1e-12
    # and both sit at rounding level of the exact sum, not merely inside the tolerance
    assert np.max(np.abs(f1[rows] - ex).max(axis=1) / fn) <= 5e-15


def test_every_tile_its_own_mass_and_some_chunks_mixed(capi):
    # tile t has mass (1 + t % 7) * 1e23: uniform tiles, different masses inside one chunk; an odd mass
    # makes chunk 10 mixed; a dead body (chunk 3) and the padding of the tail tile (last chunk) are
    # parked far away and match any mass, so their chunks stay on the uniform pass
    b = clouds.uniform_cube(N, 1500.0, 1.2, 1e24, vmax=1e8, seed=6)
    b.mass[:] = (1 + (np.arange(N) // 256) % 7) * 1e23
    dead = chunk_start(capi, 3) + 17
    b.flags[dead] &= ~np.uint8(F_EXISTS)
    b.mass[chunk_start(capi, 10) + 300] *= 3.0
    f0, p0, _, _ = one_step(capi, b, False)
    f1, p1, _, _ = one_step(capi, b, True)
    assert np.array_equal(p0, p1)
    rows = np.concatenate([np.arange(dead - 17, dead + 15), np.arange(N - 32, N)])
    for r in (rows[:32], rows[32:]):
        ex, fn = exact_rows(b, r)
        live = fn > 0
        assert np.max((np.abs(f1[r] - ex).max(axis=1) / np.where(live, fn, 1.0))[live]) <= 1e-12
    scale = np.abs(f0).max(axis=1)
    live = scale > 0
    assert np.max(np.abs(f1 - f0).max(axis=1)[live] / scale[live]) <= 1e-11
    assert np.all(f1[dead] == 0) and np.all(f0[dead] == 0)   # Compute does not run for a dead body


def test_no_uniform_tile_means_bit_identical_results(capi):
    b = clouds.uniform_cube(N, 1500.0, 1.2, 1e24, vmax=1e8, seed=7)
    b.mass[::256] *= 1.0 + 2.0 ** -40     # one odd body per tile
    f0, p0, _, _ = one_step(capi, b, False)
    f1, p1, _, _ = one_step(capi, b, True)
    assert np.array_equal(p0, p1)
    assert np.array_equal(f0.view(np.uint64), f1.view(np.uint64))


def test_uniform_pass_is_independent_of_launch_shape(capi):
    b = clouds.uniform_cube(N, 1500.0, 1.2, 1e24, vmax=1e8, seed=8)
    outs = []
    for R in ("1", "2", "4"):
        os.environ["NB_FORCE_R"] = R
        try:
            outs.append(one_step(capi, b, True))
        finally:
            os.environ.pop("NB_FORCE_R")
    for f, p, _, _ in outs[1:]:
        assert np.array_equal(f.view(np.uint64), outs[0][0].view(np.uint64)) and np.array_equal(p, outs[0][1])
