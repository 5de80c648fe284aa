"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): the i-sharded step with NCCL
exchanges must reproduce the single-GPU step bit for bit — forces, pair list, state."""
import numpy as np
import pytest

from nbodygo_b200 import clouds

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _cloud(n, mixed):
    b = clouds.uniform_cube(n, 70.0, 1.6, 1e12, vmax=50.0, seed=77)
    if mixed:  # subsume chains and fragment decisions: the whole event queue is resolved on every rank
        from nbodygo_b200.bodies import FRAGMENT, NONE, SUBSUME
        rng = np.random.default_rng(5)
        b.radius[:] = rng.uniform(0.4, 5.0, n)
        b.mass[:] = rng.uniform(1e11, 1e13, n)
        b.behavior[rng.random(n) < 0.25] = SUBSUME
        b.behavior[rng.random(n) < 0.15] = FRAGMENT
        b.behavior[rng.random(n) < 0.05] = NONE
        b.frag_factor[:] = 0.05
        b.frag_step[:] = 100.0
    return b


def _events(sim):
    return sorted((int(e["kind"]), int(e["a"]), int(e["b"]), int(e["applied"]), float(e["dist"]), float(e["f1"]))
                  for e in sim.host_events())


def _rank_main(rank, world, n, mixed, steps, q_uid, q_out):
    from nbodygo_b200 import capi
    b = _cloud(n, mixed)
    sim = capi.Sim(b.n, device=rank)
    sim.upload(b)
    if rank == 0:
        uid = capi.comm_unique_id()
        for _ in range(world - 1):
            q_uid.put(uid)
    else:
        uid = q_uid.get(timeout=120)
    sim.comm_init(rank, world, uid)
    out = []
    for _ in range(steps):
        res = sim.step(1e-3, 0.9)
        i0, i1 = sim.shard_range()
        fx, fy, fz = sim.forces()
        st = sim.download()
        out.append(dict(i0=i0, i1=i1, fx=fx[i0:i1].copy(), fy=fy[i0:i1].copy(), fz=fz[i0:i1].copy(),
                        pairs=sim.pairs(), x=st.x, vx=st.vx, vz=st.vz, flags=st.flags, rest=st.rest,
                        mass=st.mass, behavior=st.behavior, events=_events(sim), subsumed=res.n_subsumed,
                        n_pairs=res.n_pairs, n_dead=res.n_dead, resolved=res.n_resolved))
    q_out.put((rank, out))
    sim.close()


@pytest.mark.parametrize("peer_push", ["1", "0"], ids=["peer-push", "nccl-allgather"])
@pytest.mark.parametrize("world,n,mixed", [(2, 5001, False), (2, 300, False), (2, 2501, True), (4, 4001, False),
                                           (8, 9001, False), (8, 1000, False), (8, 3001, True)])
def test_sharded_step_equals_single_gpu(world, n, mixed, peer_push, monkeypatch):
    if _ndev() < world:
        pytest.skip(f"needs {world} GPUs")
    monkeypatch.setenv("NB_PEER_PUSH", peer_push)  # inherited by the spawned ranks
    import torch.multiprocessing as mp
    from nbodygo_b200 import capi
    steps = 3
    ctx = mp.get_context("spawn")
    q_uid, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, n, mixed, steps, q_uid, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q_out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    b = _cloud(n, mixed)
    sim = capi.Sim(b.n)
    sim.upload(b)
    swallowed = 0
    for k in range(steps):
        res = sim.step(1e-3, 0.9)
        fx, fy, fz = sim.forces()
        st = sim.download()
        pairs = sim.pairs()
        events = _events(sim)
        swallowed += res.n_subsumed
        assert len(pairs) > 0
        for r in range(world):
            o = got[r][k]
            sl = slice(o["i0"], o["i1"])
            assert np.array_equal(o["fx"].view(np.uint64), fx[sl].view(np.uint64))
            assert np.array_equal(o["fy"].view(np.uint64), fy[sl].view(np.uint64))
            assert np.array_equal(o["fz"].view(np.uint64), fz[sl].view(np.uint64))
            assert np.array_equal(o["pairs"], pairs)          # every rank holds the full, ordered list
            assert np.array_equal(o["x"].view(np.uint64), st.x.view(np.uint64))
            assert np.array_equal(o["vx"].view(np.uint64), st.vx.view(np.uint64))
            assert np.array_equal(o["vz"].view(np.uint64), st.vz.view(np.uint64))
            assert np.array_equal(o["flags"], st.flags) and np.array_equal(o["rest"], st.rest)
            assert o["n_pairs"] == res.n_pairs and o["n_dead"] == res.n_dead and o["resolved"] == res.n_resolved
            # ProcessMods is replicated: masses, behaviours and the event records agree on every rank
            assert np.array_equal(o["mass"].view(np.uint64), st.mass.view(np.uint64))
            assert np.array_equal(o["behavior"], st.behavior)
            assert o["events"] == events and o["subsumed"] == res.n_subsumed
    assert not mixed or swallowed > 5
    sim.close()


def _rank_sync_ops(rank, world, n, q_uid, q_out):
    """Cycle-top state sync on every rank: deletes (SetNotExists) + compaction + appends between steps."""
    from nbodygo_b200 import capi
    from nbodygo_b200.bodies import BodyArrays
    b = clouds.uniform_cube(n, 90.0, 1.6, 1e12, vmax=50.0, seed=78)
    sim = capi.Sim(b.n + 64, device=rank)
    sim.upload(b)
    if rank == 0:
        uid = capi.comm_unique_id()
        for _ in range(world - 1):
            q_uid.put(uid)
    else:
        uid = q_uid.get(timeout=120)
    sim.comm_init(rank, world, uid)
    sim.step(1e-3, 1.0)
    for i in (7, n // 2, n - 1):
        sim.patch(i, 1, mass=np.zeros(1), flags=np.zeros(1, dtype=np.uint8))
    new_n, _ = sim.compact()
    add = clouds.uniform_cube(5, 20.0, 1.0, 1e12, seed=3)
    sim.append(add, R=0.5)
    res = sim.step(1e-3, 0.5)
    st = sim.download()
    q_out.put((rank, dict(new_n=new_n, n=sim.count(), x=st.x, vx=st.vx, rest=st.rest, pairs=sim.pairs(),
                          n_pairs=res.n_pairs, shard=sim.shard_range())))
    sim.close()


@pytest.mark.parametrize("peer_push", ["1", "0"], ids=["peer-push", "nccl-allgather"])
def test_state_sync_ops_on_two_gpus(peer_push, monkeypatch):
    world, n = 2, 3001
    if _ndev() < world:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("NB_PEER_PUSH", peer_push)
    import torch.multiprocessing as mp
    from nbodygo_b200 import capi
    ctx = mp.get_context("spawn")
    q_uid, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_sync_ops, args=(r, world, n, q_uid, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q_out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = clouds.uniform_cube(n, 90.0, 1.6, 1e12, vmax=50.0, seed=78)
    sim = capi.Sim(b.n + 64)
    sim.upload(b)
    sim.step(1e-3, 1.0)
    for i in (7, n // 2, n - 1):
        sim.patch(i, 1, mass=np.zeros(1), flags=np.zeros(1, dtype=np.uint8))
    new_n, _ = sim.compact()
    sim.append(clouds.uniform_cube(5, 20.0, 1.0, 1e12, seed=3), R=0.5)
    res = sim.step(1e-3, 0.5)
    st = sim.download()
    for r in range(world):
        o = got[r]
        assert o["new_n"] == new_n == n - 3 and o["n"] == n + 2
        assert np.array_equal(o["x"].view(np.uint64), st.x.view(np.uint64))
        assert np.array_equal(o["vx"].view(np.uint64), st.vx.view(np.uint64))
        assert np.array_equal(o["rest"], st.rest) and np.array_equal(o["pairs"], sim.pairs())
        assert o["n_pairs"] == res.n_pairs
    assert got[0]["shard"][1] == got[1]["shard"][0]
    sim.close()
