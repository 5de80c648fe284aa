#!/usr/bin/env python3
"""NVLink bytes of the fused state exchange, per kernel launch, from ncu — ONE process driving P handles
(one per GPU, one thread each), which is the configuration ncu can profile (never a multi-rank command).

  ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum --clock-control none \
      --kernel-name-base demangled -k regex:'k_integrate|k_push_pairs|k_push_shard' --csv \
      --log-file gpurun_out/nvlink_ncu.csv python tools/nvlink_ncu.py --gpus 2 --config C4 --n 200000 --cycles 2

Expected per k_integrate launch on P GPUs: 57 B x ceil(n/P) x (P-1) transmitted (x y z vx vy vz rest = 7 x 8 B
+ 1 flag byte per body of the shard, stored into each of the P-1 peers).  Without ncu the script just runs the
cycles and prints the exchange mode.  `nvidia-smi nvlink -gt d` reports N/A in this container
(tools/nvlink_bytes.py), hence ncu."""
import argparse
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--config", default="C4")
    ap.add_argument("--n", type=int, default=200_000)
    ap.add_argument("--cycles", type=int, default=2)
    a = ap.parse_args()
    from nbodygo_b200 import capi, clouds
    b = clouds.config(a.config, n=a.n)
    P = a.gpus
    sims = [capi.Sim(b.n, device=r) for r in range(P)]
    for s in sims:
        s.upload(b)
    uid = capi.comm_unique_id()
    errs = []

    def join(r):   # ncclCommInitRank blocks until every rank has joined: one thread per handle
        try:
            sims[r].comm_init(r, P, uid)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=join, args=(r,)) for r in range(P)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    print("exchange:", [capi.COMM_MODE_NAMES.get(s.comm_mode()) for s in sims], flush=True)
    out = [None] * P

    def run(r):    # one host thread per handle, like one goroutine per handle in the Go host
        try:
            log = []
            for _ in range(a.cycles):
                res = sims[r].step(1e-9, 1.0)
                log.append((res.n_pairs, round(res.ms_total, 3)))
            out[r] = log
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=run, args=(r,)) for r in range(P)]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]
    for r in range(P):
        print(f"handle {r}: (pairs, ms_total) per cycle = {out[r]}", flush=True)
    for s in sims:
        s.close()


if __name__ == "__main__":
    main()
