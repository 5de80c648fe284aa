#!/usr/bin/env python3
"""Cycle time of small collections with and without the CUDA-graph replay (NB_GRAPH=1/0):
device time (CUDA events around the cycle) and wall time per nb_step call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbodygo_b200 import capi, clouds  # noqa: E402

for n in (1000, 3000, 10000, 32000):
    b = clouds.config("C4", n=n)
    for g in ("0", "1"):
        os.environ["NB_GRAPH"] = g
        sim = capi.Sim(n)
        sim.upload(b)
        for _ in range(20):
            sim.step(1e-9, 1.0)
        steps = 300
        t0 = time.perf_counter()
        dev = 0.0
        for _ in range(steps):
            dev += sim.step(1e-9, 1.0).ms_total
        wall = (time.perf_counter() - t0) / steps
        print(f"n={n:6d} graph={g}: device {1e3 * dev / steps:8.1f} us/cycle  wall {1e6 * wall:8.1f} us/step  "
              f"graph stats {sim.graph_stats()}", flush=True)
        sim.close()
