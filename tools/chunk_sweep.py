#!/usr/bin/env python3
"""Development sweep of the K1 launch parameters at mid sizes: chunk count (NB_CHUNKS) x launch
shape (NB_FORCE_R) -> K1 time and whole-cycle time.  The production choice (chunking() and
launch_force()) is the row marked '*'.

  python tools/chunk_sweep.py --n 10000,32000,100000"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", default="10000,32000,100000")
    ap.add_argument("--chunks", default="0,16,24,32,48,64,96,128")
    ap.add_argument("--codes", default="0,4,2,1")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tj", default="default,large", help="tile size: default rule, or 256-body tiles everywhere")
    a = ap.parse_args()
    from nbodygo_b200 import capi, clouds
    peak, _ = capi.measure_fp64_peak(0, 4096)
    print(f"fp64 peak measured {peak:.2f} TFLOP/s")
    for n in (int(v) for v in a.n.split(",")):
        b = clouds.config("C4", n=n)
        for tj, code in ((t, c) for t in a.tj.split(",") for c in a.codes.split(",")):
            if tj == "large":
                os.environ["NB_TJ_SMALL_BELOW"] = "0"
            else:
                os.environ.pop("NB_TJ_SMALL_BELOW", None)
            os.environ["NB_FORCE_R"] = code
            sim = capi.Sim(b.n)   # NB_FORCE_R is read at nb_create
            sim.upload(b)
            for ch in a.chunks.split(","):
                if ch == "0":
                    os.environ.pop("NB_CHUNKS", None)
                else:
                    os.environ["NB_CHUNKS"] = ch   # read by every nb_step
                sim.step(1e-9, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE | capi.STEP_PHASE_TIMINGS)
                rs = [sim.step(1e-9, 1.0, capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE | capi.STEP_PHASE_TIMINGS) for _ in range(a.reps)]
                k1 = min(r.ms_force for r in rs)
                tot = min(r.ms_total for r in rs)
                frac = 30.0 * n * (n - 1.0) / (k1 * 1e-3) / 1e12 / peak
                star = "*" if (code == "0" and ch == "0" and tj == "default") else " "
                print(f"{star} n={n:7d} tj={tj:>7} shape={code:>4} chunks={ch:>3}: K1 {k1 * 1e3:9.1f} us  cycle {tot * 1e3:9.1f} us"
                      f"  K1 {100 * frac:5.1f}% of peak", flush=True)
            sim.close()
    for k in ("NB_CHUNKS", "NB_FORCE_R", "NB_TJ_SMALL_BELOW"):
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()
