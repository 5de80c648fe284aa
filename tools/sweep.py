#!/usr/bin/env python3
"""N sweep (BASELINE config C5): interactions/s and steps/s vs body count, 1..8 GPUs.

  python tools/sweep.py --ns 1000,2000,...                      # one GPU
  python -m torch.distributed.run --nproc-per-node 8 ... tools/sweep.py --ns ...

Same law as C4 (uniform sphere, elastic, ~1e-3*n overlapping pairs), CUDA-event timing of
nb_step (max over ranks).  Prints one JSON line per n (rank 0)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ns", default="1000,2000,4000,8000,16000,32000,64000,128000,256000,512000,1000000")
    ap.add_argument("--budget-s", type=float, default=4.0, help="target timed seconds per n")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from nbodygo_b200 import capi, clouds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peak, _ = capi.measure_fp64_peak(local, 2048)
    for n in [int(s) for s in a.ns.split(",")]:
        b = clouds.config("C4", n=n)
        sim = capi.Sim(n, device=local)
        sim.upload(b)
        if world > 1:
            uid = [capi.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            sim.comm_init(rank, world, uid[0])
        for _ in range(3):
            r = sim.step(1e-9, 1.0, capi.STEP_DEFAULT | capi.STEP_PHASE_TIMINGS)
        steps = int(max(3, min(2000, a.budget_s / max(r.ms_total * 1e-3, 1e-5))))
        if world > 1:
            t = torch.tensor([steps], device="cuda"); dist.broadcast(t, 0); steps = int(t.item())
            dist.barrier()
        torch.cuda.synchronize()
        ms, ms_force, pairs = 0.0, 0.0, 0
        import time
        t0 = time.perf_counter()
        for _ in range(steps):
            r = sim.step(1e-9, 1.0, capi.STEP_DEFAULT | capi.STEP_PHASE_TIMINGS)
            ms += r.ms_total; ms_force += r.ms_force; pairs += r.n_pairs
        wall = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([ms, ms_force, wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_force, wall = (float(v) for v in t)
        if rank == 0:
            inter = float(n) * (n - 1.0)
            print(json.dumps({
                "n": n, "n_gpus": world, "steps": steps, "ms_per_step_device": ms / steps,
                "ms_per_step_wall": 1e3 * wall / steps, "steps_per_s_wall": steps / wall,
                "interactions_per_s": inter * steps / (ms * 1e-3),
                "frac_of_measured_fp64_peak": 30 * inter * steps / (ms * 1e-3) / 1e12 / (peak * world),
                "k_force_ms": ms_force / steps, "pairs_per_step": pairs / steps}), flush=True)
        sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
