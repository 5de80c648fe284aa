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


# ------------------------------------------------------------------------------------------------
# The path the scaling bench times: n >= 16,384 (256-body tiles), the uniform-mass / per-body-mass
# split of K1, launch shapes that differ between the shard and the whole array (R is chosen from the
# shard size).  Arrays of these sizes are compared through digests computed on each rank.
def _digest(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _big_cloud(kind, n):
    if kind == "C3":          # BASELINE config 3: 100,000-body cube cloud, elastic (one mass: uniform pass)
        return clouds.config("C3", n=n)
    if kind == "C4":          # BASELINE config 4
        return clouds.config("C4", n=n)
    if kind == "mixed-mass":  # per-body masses: every chunk on the general pass
        b = clouds.uniform_cube(n, 900.0, 1.9, 1e12, vmax=50.0, seed=81)
        b.mass[:] = np.random.default_rng(9).uniform(1e11, 1e13, n)
        return b
    raise ValueError(kind)


def _big_state_digests(sim, i0, i1):
    fx, fy, fz = sim.forces()
    st = sim.download()
    return dict(force=_digest(fx[i0:i1], fy[i0:i1], fz[i0:i1]), pairs=_digest(sim.pairs()),
                state=_digest(st.x, st.y, st.z, st.vx, st.vy, st.vz, st.rest, st.flags, st.mass))


def _rank_big(rank, world, kind, n, steps, shard_upload, q_uid, q_out):
    from nbodygo_b200 import capi
    b = _big_cloud(kind, n)
    sim = capi.Sim(b.n, device=rank)
    if rank == 0:
        uid = capi.comm_unique_id()
        for _ in range(world - 1):
            q_uid.put(uid)
    else:
        uid = q_uid.get(timeout=180)
    if shard_upload:   # each rank hands over only its slice; the rest arrives over NVLink
        sim.comm_init(rank, world, uid)
        i0, i1, _, _ = capi.plan(b.n, rank, world)
        sl = slice(i0, i1)
        sim.upload_shard(b.n, i0, i1 - i0, b.x[sl], b.y[sl], b.z[sl], b.vx[sl], b.vy[sl], b.vz[sl], b.mass[sl],
                         b.radius[sl], rest=b.rest[sl], ff=b.frag_factor[sl], fs=b.frag_step[sl],
                         behavior=b.behavior[sl], flags=b.flags[sl])
    else:
        sim.upload(b)
        sim.comm_init(rank, world, uid)
    out = []
    for _ in range(steps):
        res = sim.step(1e-9 if kind in ("C3", "C4") else 1e-3, 0.9)
        i0, i1 = sim.shard_range()
        d = _big_state_digests(sim, i0, i1)
        d.update(i0=i0, i1=i1, n_pairs=res.n_pairs, resolved=res.n_resolved, mode=sim.comm_mode())
        out.append(d)
    q_out.put((rank, out))
    sim.close()


@pytest.mark.parametrize("peer_push", ["1", "0"], ids=["peer-push", "nccl-allgather"])
@pytest.mark.parametrize("world,kind,n,shard_upload", [
    (2, "C4", 40_000, False), (2, "mixed-mass", 40_000, False), (2, "C3", 100_000, False),
    (2, "C3", 100_000, True), (4, "C4", 200_000, True), (8, "C4", 1_000_000, True)],
    ids=["2xC4-40k", "2xmixed-40k", "2xC3-100k", "2xC3-100k-shard-upload", "4xC4-200k", "8xC4-1M"])
def test_production_path_sharded_equals_single_gpu(world, kind, n, shard_upload, peer_push, monkeypatch):
    """Sharded-vs-single-GPU on the configuration SCALE times: 256-body tiles, the FORCE_UNI / FORCE_MIXED
    split, R differing between shard and whole array; BASELINE config 3 on 2 GPUs, C4 on 8.  Forces of every
    shard, the gathered pair list and the whole state must be bit-identical to one GPU.  The slice rule is
    cmd/runner/computation-runner.go:286-305."""
    if _ndev() < world:
        pytest.skip(f"needs {world} GPUs")
    if world == 8 and peer_push == "0":
        pytest.skip("the 1M-body case runs once (peer-push, the production exchange)")
    monkeypatch.setenv("NB_PEER_PUSH", peer_push)
    import torch.multiprocessing as mp
    from nbodygo_b200 import capi
    steps = 1 if n >= 1_000_000 else 2
    ctx = mp.get_context("spawn")
    q_uid, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_big, args=(r, world, kind, n, steps, shard_upload, q_uid, q_out))
             for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q_out.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    b = _big_cloud(kind, n)
    assert b.n >= 16384
    sim = capi.Sim(b.n)
    sim.upload(b)
    assert sim.comm_mode() == capi.COMM_SINGLE
    for k in range(steps):
        res = sim.step(1e-9 if kind in ("C3", "C4") else 1e-3, 0.9)
        assert res.n_pairs > 0
        for r in range(world):
            o = got[r][k]
            assert o["mode"] == (capi.COMM_PEER_PUSH if peer_push == "1" else capi.COMM_NCCL)
            ref = _big_state_digests(sim, o["i0"], o["i1"])
            assert o["force"] == ref["force"], f"step {k} rank {r}: forces of the shard"
            assert o["pairs"] == ref["pairs"], f"step {k} rank {r}: pair list"
            assert o["state"] == ref["state"], f"step {k} rank {r}: state"
            assert o["n_pairs"] == res.n_pairs and o["resolved"] == res.n_resolved
    sim.close()


def _frag_scene(n):
    """Fragment-behaviour bodies that start fragmenting in cycle 1 (the host does not spawn anything
    here: the flag stays set, so the bodies keep applying the force of their last Compute)."""
    from nbodygo_b200.bodies import FRAGMENT
    b = clouds.uniform_cube(n, 70.0, 1.6, 1e12, vmax=500.0, seed=91)
    b.behavior[::3] = FRAGMENT
    b.frag_factor[:] = 0.05
    b.frag_step[:] = 100.0
    return b


def _frag_ops(sim, n):
    """Two cycles, deletes + compaction + appends (the shard boundary moves), two more cycles."""
    log = []
    for _ in range(2):
        sim.step(1e-3, 0.9)
    for i in range(0, n // 2, 5):   # a tenth of the bodies, all from the first rank's half: the boundary moves far
        sim.patch(i, 1, mass=np.zeros(1), flags=np.zeros(1, dtype=np.uint8))
    new_n, old = sim.compact()
    sim.append(clouds.uniform_cube(7, 20.0, 1.0, 1e12, seed=4), R=0.9)
    for _ in range(2):
        res = sim.step(1e-3, 0.9)
        st = sim.download()
        fx, fy, fz = sim.forces()
        log.append(dict(x=st.x, y=st.y, vx=st.vx, vz=st.vz, flags=st.flags, n_pairs=res.n_pairs, fx=fx, fz=fz))
    return new_n, old, log


def _rank_frag(rank, world, n, q_uid, q_out):
    from nbodygo_b200 import capi
    sim = capi.Sim(n + 64, device=rank)
    sim.upload(_frag_scene(n))
    if rank == 0:
        uid = capi.comm_unique_id()
        for _ in range(world - 1):
            q_uid.put(uid)
    else:
        uid = q_uid.get(timeout=120)
    sim.comm_init(rank, world, uid)
    new_n, _, log = _frag_ops(sim, n)
    i0, i1 = sim.shard_range()
    q_out.put((rank, dict(new_n=new_n, log=log, i0=i0, i1=i1)))
    sim.close()


@pytest.mark.parametrize("peer_push", ["1", "0"], ids=["peer-push", "nccl-allgather"])
def test_fragmenting_bodies_keep_their_force_when_the_shards_move(peer_push, monkeypatch):
    """A fragmenting body does not compute; Update keeps applying the force of its last Compute
    (body.go:152-155,119-123).  Compaction and appends move the ceil(n/P) shard boundaries, so the body
    may change owner: its stored force must be on every rank (K4 pushes it / the fallback gathers it)."""
    world, n = 2, 3000
    if _ndev() < world:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("NB_PEER_PUSH", peer_push)
    import torch.multiprocessing as mp
    from nbodygo_b200 import capi
    from nbodygo_b200.bodies import F_EXISTS, F_FRAGMENTING
    ctx = mp.get_context("spawn")
    q_uid, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_rank_frag, args=(r, world, n, q_uid, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q_out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sim = capi.Sim(n + 64)
    sim.upload(_frag_scene(n))
    new_n, old, log = _frag_ops(sim, n)
    frag = (log[0]["flags"] & F_FRAGMENTING) != 0
    live = (log[0]["flags"] & F_EXISTS) != 0
    # fragmenting bodies that were computed by rank 1 before the compaction and belong to rank 0 after it
    old_boundary, new_boundary = -(-n // world), -(-(new_n + 7) // world)
    moved = np.zeros(len(frag), dtype=bool)
    moved[:new_n] = (old >= old_boundary) & (np.arange(new_n) < new_boundary)
    assert (frag & live & moved).sum() >= 3, "the scene must move fragmenting bodies to another owner"
    for r in range(world):
        o = got[r]
        assert o["new_n"] == new_n
        for k, ref in enumerate(log):
            for f in ("x", "y", "vx", "vz"):
                assert np.array_equal(o["log"][k][f].view(np.uint64), ref[f].view(np.uint64)), (r, k, f)
            assert np.array_equal(o["log"][k]["flags"], ref["flags"]) and o["log"][k]["n_pairs"] == ref["n_pairs"]
            sl = slice(o["i0"], o["i1"])
            assert np.array_equal(o["log"][k]["fx"][sl].view(np.uint64), ref["fx"][sl].view(np.uint64))
    sim.close()
