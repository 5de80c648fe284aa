"""CPU tests of the N>1 protocol with world_size-2/3 `gloo` processes.

The sharded cycle the library runs on GPUs — plan the i-range (nb_plan), compute the own
shard against all j, all-gather the pair lists in rank order, resolve replicated, integrate
the own shard, all-gather the shard's state — is executed here with the CPU oracle as the
per-rank engine and torch.distributed/gloo as the transport, and must reproduce the
single-rank cycle bit for bit (pair list order included)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nbodygo_b200 import capi, clouds
from nbodygo_b200.bodies import BodyArrays

STATE = ("x", "y", "z", "vx", "vy", "vz", "rest")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_var(arr: np.ndarray, world: int) -> list:
    """all-gather of variable-length int32 [k,2] arrays: counts first, then padded lists."""
    cnt = torch.tensor([len(arr)], dtype=torch.int64)
    cnts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    mx = max(int(c.item()) for c in cnts)
    pad = torch.zeros((max(mx, 1), arr.shape[1]), dtype=torch.int32)
    pad[: len(arr)] = torch.from_numpy(arr.astype(np.int32))
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return [o[: int(c.item())].numpy() for o, c in zip(outs, cnts)]


def _cloud(n, mixed):
    b = clouds.uniform_cube(n, 40.0, 1.6, 1e12, vmax=50.0, seed=99)
    if mixed:  # subsume chains + fragment decisions travel in the same event list
        from nbodygo_b200.bodies import FRAGMENT, SUBSUME
        rng = np.random.default_rng(17)
        b.radius[:] = rng.uniform(0.4, 5.0, n)
        b.mass[:] = rng.uniform(1e11, 1e13, n)
        b.behavior[rng.random(n) < 0.3] = SUBSUME
        b.behavior[rng.random(n) < 0.15] = FRAGMENT
        b.frag_factor[:] = 0.05
        b.frag_step[:] = 100.0
    return b


def _worker(rank, world, port, n, mixed, steps, q):
    from oracle.oracle import EVENT_DTYPE, EV_COLLISION, OracleSim
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = _cloud(n, mixed)
    o = OracleSim(b)
    pair_log = []
    for _ in range(steps):
        i0, i1, _, _ = capi.plan(b.n, rank, world)
        o.compute(i0, i1)                                   # K1 on the own shard
        # the shard's event list (collisions and subsumes, arrival order) as (kind, a, b) triples —
        # what k_push_pairs / the NCCL all-gather move between the GPUs
        mine = np.stack([o.events["kind"], o.events["a"], o.events["b"]], axis=1).astype(np.int32)
        alle = np.concatenate(_gather_var(mine.reshape(-1, 3), world))   # rank order == i order
        pair_log.append(alle[alle[:, 0] == EV_COLLISION][:, 1:].copy())
        ev = np.zeros(len(alle), dtype=EVENT_DTYPE)
        ev["kind"], ev["a"], ev["b"] = alle[:, 0], alle[:, 1], alle[:, 2]
        o.process_mods(ev)                                  # K3 replicated on every rank (mass, Exists too)
        o.update(1e-3, 0.9, i0, i1)                         # K4 on the own shard
        shard = (b.n + world - 1) // world
        for f in STATE:                                     # state exchange (padded equal shards)
            a = getattr(b, f)
            send = torch.zeros(shard, dtype=torch.float64)
            send[: i1 - i0] = torch.from_numpy(a[i0:i1])
            outs = [torch.zeros(shard, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(outs, send)
            a[:] = torch.cat(outs)[: b.n].numpy()
        fl = torch.zeros(shard, dtype=torch.uint8)
        fl[: i1 - i0] = torch.from_numpy(b.flags[i0:i1])
        outs = [torch.zeros(shard, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(outs, fl)
        b.flags[:] = torch.cat(outs)[: b.n].numpy()
    if rank == world - 1:
        q.put((b.x.copy(), b.vx.copy(), b.vz.copy(), b.flags.copy(), b.mass.copy(), b.behavior.copy(),
               [p.tolist() for p in pair_log]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,mixed", [(2, 301, False), (3, 200, False), (2, 400, True)])
def test_sharded_cycle_equals_single_rank(world, n, mixed):
    from oracle.oracle import OracleSim
    steps = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, mixed, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    x, vx, vz, flags, mass, behavior, pair_log = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    b = _cloud(n, mixed)
    m_before = b.mass.copy()
    o = OracleSim(b)
    for k in range(steps):
        o.compute()
        assert o.collision_pairs().tolist() == pair_log[k], f"pair list differs at step {k}"
        o.process_mods()
        o.update(1e-3, 0.9)
    assert len(pair_log[0]) > 10
    assert np.array_equal(x.view(np.uint64), b.x.view(np.uint64))
    assert np.array_equal(vx.view(np.uint64), b.vx.view(np.uint64))
    assert np.array_equal(vz.view(np.uint64), b.vz.view(np.uint64))
    assert np.array_equal(flags, b.flags)
    assert np.array_equal(mass.view(np.uint64), b.mass.view(np.uint64)) and np.array_equal(behavior, b.behavior)
    assert not mixed or ((m_before != b.mass).sum() > 10 and (~b.exists).sum() > 5)


def test_plan_is_a_partition_and_depends_on_n_only():
    for n in (0, 1, 255, 256, 257, 1000, 10_000, 1_000_000, 4_000_000):
        for world in (1, 2, 3, 4, 8):
            ranges = [capi.plan(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            shard = -(-n // world) if n else 0
            assert all(i1 - i0 <= shard for i0, i1, _, _ in ranges)
            # the j-chunking (summation order) must not depend on the rank count
            assert len({(nc, tpc) for _, _, nc, tpc in ranges}) == 1
            assert ranges[0][2:] == capi.plan(n, 0, 1)[2:]
        _, _, nc, tpc = capi.plan(n, 0, 1)
        tj = 64 if n < 16384 else (512 if n >= 786432 else 256)     # the tile size is a function of n only, too
        tiles = -(-n // tj)
        assert 1 <= nc <= 128 and nc * tpc >= tiles
    with pytest.raises(capi.NbError):
        capi.plan(10, 2, 2)
