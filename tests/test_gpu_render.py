"""Render path: the float32 Renderable snapshot of every cycle lands in
library-owned pinned host buffers as part of the cycle's own stream (nb_render_buffers)."""
import numpy as np
import pytest

from nbodygo_b200 import clouds

pytestmark = pytest.mark.gpu


def test_pinned_render_buffers_follow_every_step():
    from nbodygo_b200 import capi
    b = clouds.uniform_cube(3000, 100.0, 1.0, 1e13, vmax=30.0, seed=71)
    b.flags[[4, 2000]] = 0
    sim = capi.Sim(b.n + 8)
    sim.upload(b)
    xyz, ex = sim.render_buffers()
    for _ in range(3):
        sim.step(1e-2, 1.0)
        st = sim.download()
        assert np.array_equal(ex[: b.n], st.exists.astype(np.uint8))
        live = st.exists
        assert np.array_equal(xyz[: b.n][live, 0], st.x[live].astype(np.float32))
        assert np.array_equal(xyz[: b.n][live, 2], st.z[live].astype(np.float32))
        assert np.all(xyz[: b.n][~live] == 0)
        x2, e2 = sim.render()                       # the explicit-copy call returns the same snapshot
        assert np.array_equal(x2, xyz[: b.n]) and np.array_equal(e2, ex[: b.n])
    sim.close()


def _rank_render(rank, world, n, q_uid, q_out):
    from nbodygo_b200 import capi
    b = clouds.uniform_cube(n, 90.0, 1.6, 1e12, vmax=50.0, seed=79)
    sim = capi.Sim(b.n, device=rank)
    sim.upload(b)
    if rank == 0:
        uid = capi.comm_unique_id()
        for _ in range(world - 1):
            q_uid.put(uid)
    else:
        uid = q_uid.get(timeout=120)
    sim.comm_init(rank, world, uid)
    xyz, ex = sim.render_buffers()
    sim.step(1e-3, 1.0)
    st = sim.download()
    q_out.put((rank, xyz[:n].copy(), ex[:n].copy(), st.x.copy()))
    sim.close()


def test_render_snapshot_covers_all_bodies_on_every_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, n = 2, 2001
    ctx = mp.get_context("spawn")
    q_uid, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_render, args=(r, world, n, q_uid, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q_out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, xyz, ex, x in got:
        assert ex.all()
        assert np.array_equal(xyz[:, 0], x.astype(np.float32))   # every body, not only the rank's own shard
    assert np.array_equal(got[0][1], got[1][1])
