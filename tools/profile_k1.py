#!/usr/bin/env python3
"""Workload for the ncu captures of K1: uploads a BASELINE cloud and runs a few force+detect
cycles (no integrate, so every cycle sees the same state).  Meant to be run under ncu:

  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:'k_force<.*, 2>' --launch-skip 1 -c 1 -o gpurun_out/k_force_uni \
      python tools/profile_k1.py --config C4

Development tool; numbers printed under a profiler are never bench values."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--cycles", type=int, default=2)
    ap.add_argument("--integrate", action="store_true", help="full cycles (K3 + K4 too)")
    a = ap.parse_args()
    from nbodygo_b200 import capi, clouds
    b = clouds.config(a.config, n=a.n or None)
    sim = capi.Sim(b.n)
    sim.upload(b)
    opts = capi.STEP_COLLISIONS | capi.STEP_PHASE_TIMINGS | (0 if a.integrate else capi.STEP_NO_INTEGRATE)
    for _ in range(a.cycles):
        r = sim.step(1e-9, 1.0, opts)
        print(f"n={b.n} ms_force={r.ms_force:.3f} pairs={r.n_pairs}", flush=True)
    sim.close()


if __name__ == "__main__":
    main()
