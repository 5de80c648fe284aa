"""The cycle as a CUDA graph (single GPU): a cycle whose parameters repeat is captured once and
replayed with one launch.  Replays must be indistinguishable from plain stream launches — state
bits, pair lists, counters, phase timings — and every change of the cycle's parameters (body count,
array pointers after a compaction, time scaling, options, render buffers) must invalidate the graph."""
import numpy as np
import pytest

from nbodygo_b200 import clouds
from nbodygo_b200.bodies import SUBSUME

pytestmark = pytest.mark.gpu


def _cloud(n=3000):
    b = clouds.uniform_cube(n, 60.0, 1.5, 1e12, vmax=80.0, seed=31)
    b.behavior[::7] = SUBSUME
    b.radius[::7] *= 2.0
    return b


def _run(monkeypatch, graph, steps=8):
    from nbodygo_b200 import capi
    monkeypatch.setenv("NB_GRAPH", "1" if graph else "0")
    sim = capi.Sim(4096)
    sim.upload(_cloud())
    out = []
    for _ in range(steps):
        r = sim.step(1e-4, 0.9, capi.STEP_DEFAULT | capi.STEP_PHASE_TIMINGS)
        st = sim.download()
        out.append((r.n_pairs, r.n_resolved, r.n_subsumed, r.n_dead, r.n_host_events, sim.pairs().copy(),
                    st.x.copy(), st.vx.copy(), st.mass.copy(), st.flags.copy(), r.ms_total, r.ms_force))
    stats = sim.graph_stats()
    launches = sim.launch_count()
    sim.close()
    return out, stats, launches


def test_graph_replay_is_bit_identical_to_stream_launches(monkeypatch):
    a, sa, la = _run(monkeypatch, graph=True)
    b, sb, lb = _run(monkeypatch, graph=False)
    assert sb == (0, 0)
    assert sa[0] == 1 and sa[1] == 6, sa          # step 1 plain, step 2 captures (+runs), steps 3..8 replay
    assert la == lb                                 # the launch count is that of the kernels, graph or not
    assert sum(s[0] for s in a) > 0 and sum(s[2] for s in a) > 0
    for k, (ga, gb) in enumerate(zip(a, b)):
        assert ga[:5] == gb[:5], f"step {k}: counters"
        assert np.array_equal(ga[5], gb[5]), f"step {k}: pair list"
        for u, v in zip(ga[6:9], gb[6:9]):
            assert np.array_equal(u.view(np.uint64), v.view(np.uint64)), f"step {k}: state bits"
        assert np.array_equal(ga[9], gb[9])
        assert ga[10] > 0 and ga[11] > 0 and ga[10] >= ga[11]   # phase timings come from event nodes of the graph


def test_graph_is_invalidated_by_every_parameter_change(monkeypatch):
    from nbodygo_b200 import capi
    monkeypatch.setenv("NB_GRAPH", "1")
    b = _cloud(2000)
    ref = capi.Sim(4096)
    monkeypatch.setenv("NB_GRAPH", "0")
    plain = capi.Sim(4096)
    for s in (ref, plain):
        s.upload(b)

    def both(fn):
        return fn(ref), fn(plain)

    def same_state():
        x, y = ref.download(), plain.download()
        assert x.n == y.n
        for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "rest"):
            assert np.array_equal(getattr(x, f).view(np.uint64), getattr(y, f).view(np.uint64)), f
        assert np.array_equal(x.flags, y.flags)

    for _ in range(4):
        both(lambda s: s.step(1e-4, 0.9))
    same_state()
    c0, r0 = ref.graph_stats()
    assert c0 == 1 and r0 == 2
    # new data under the same parameters: the graph stays valid and sees the patched values
    both(lambda s: s.patch(5, 1, vx=np.array([123.0])))
    both(lambda s: s.step(1e-4, 0.9))
    same_state()
    assert ref.graph_stats() == (1, 3)
    # time scaling, R and the options are kernel parameters
    for ts, R, opts in ((2e-4, 0.9, capi.STEP_DEFAULT), (2e-4, 0.5, capi.STEP_DEFAULT), (2e-4, 0.5, 0)):
        for _ in range(3):
            both(lambda s: s.step(ts, R, opts))
        same_state()
    assert ref.graph_stats()[0] == 4
    # compaction swaps array pointers and changes n; append changes n
    both(lambda s: s.compact())
    for _ in range(3):
        both(lambda s: s.step(1e-4, 0.9))
    same_state()
    both(lambda s: s.append(clouds.uniform_cube(9, 10.0, 1.0, 1e12, seed=2), R=0.9))
    for _ in range(3):
        both(lambda s: s.step(1e-4, 0.9))
    same_state()
    # render buffers requested after a graph exists
    xyz, ex = ref.render_buffers()
    for _ in range(3):
        r1, r2 = both(lambda s: s.step(1e-4, 0.9))
    same_state()
    px, pe = plain.render()
    n = ref.count()
    assert np.array_equal(xyz[:n], px.reshape(-1, 3)[:n]) and np.array_equal(ex[:n], pe[:n])
    assert plain.graph_stats() == (0, 0)
    ref.close()
    plain.close()
