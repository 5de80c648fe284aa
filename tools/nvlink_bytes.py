#!/usr/bin/env python3
"""NVLink traffic of the fused exchange: cumulative per-link data counters (`nvidia-smi nvlink -gt d`)
before and after K cycles of the sharded step, per GPU, next to what the design predicts.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/nvlink_bytes.py --config C4 --cycles 20 > gpurun_out/nvlink.json

Per cycle and rank the exchange stores the rank's shard of x y z vx vy vz rest (7 x 8 B) + flags (1 B) = 57 B/body
into each of the P-1 peers (K4), plus the pair list (8 B/event, tens of KB): expected tx per rank and cycle
= 57 B x ceil(n/P) x (P-1).  Development tool: the counters are whole-GPU (NCCL's communicator set-up and the
uploads are kept outside the measured window)."""
import argparse
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def read_counters():
    """{gpu index: (tx KiB, rx KiB)} summed over links."""
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d"], capture_output=True, text=True).stdout
    res, gpu = {}, None
    for line in out.splitlines():
        m = re.match(r"GPU (\d+):", line)
        if m:
            gpu = int(m.group(1))
            res[gpu] = [0, 0]
            continue
        m = re.search(r"Link \d+: Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None:
            res[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2))
    return res, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--cycles", type=int, default=20)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from nbodygo_b200 import capi, clouds
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b = clouds.config(a.config, n=a.n or None)
    sim = capi.Sim(b.n, device=local)
    sim.upload(b)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim.comm_init(rank, world, uid[0])
    for _ in range(2):
        sim.step(1e-9, 1.0)
    dist.barrier(); torch.cuda.synchronize()
    before, raw0 = read_counters() if rank == 0 else (None, None)
    dist.barrier()
    pairs = 0
    for _ in range(a.cycles):
        pairs += sim.step(1e-9, 1.0).n_pairs
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0:
        after, raw1 = read_counters()
        shard = -(-b.n // world)
        exp = 57 * shard * (world - 1)
        per_gpu = {g: {"tx_bytes_per_cycle": (after[g][0] - before[g][0]) * 1024 / a.cycles,
                       "rx_bytes_per_cycle": (after[g][1] - before[g][1]) * 1024 / a.cycles} for g in sorted(after)
                   if g in before}
        print(json.dumps({"config": a.config, "n_bodies": b.n, "n_gpus": world, "cycles": a.cycles,
                          "exchange": capi.COMM_MODE_NAMES.get(sim.comm_mode()),
                          "expected_tx_bytes_per_rank_per_cycle": exp,
                          "expected_note": "57 B x ceil(n/P) x (P-1) state + 8 B x events x (P-1) pair list",
                          "pairs_per_cycle": pairs / a.cycles, "per_gpu": per_gpu,
                          "raw_sample": raw1.splitlines()[:6]}))
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
