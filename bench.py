#!/usr/bin/env python3
"""bench.py — headline benchmark of the nbodygo hot path on B200.

Metric (BASELINE.json): fp64 body-pair interactions/s (and steps/s) of one full
compute cycle — all-pairs force + collision detect + elastic resolve + integrate —
on config C4: 1,000,000-body uniform sphere with elastic collisions, i-sharded over
N GPUs (strong scaling: the body count is fixed).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      # CPU work-pool port

One JSON line on stdout (rank 0).  `value` is device-resident throughput (CUDA
events around nb_step, max over ranks); `e2e` goes through the C ABI with HOST
buffers every step (upload → step → download inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_INTERACTION = 30.0          # README.md:20,25 of the reference
NOMINAL_FP64_TFLOPS = 37.2            # 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4")
    ap.add_argument("--bodies", dest="n", type=int, default=0, help="override the body count (development only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--time-scaling", type=float, default=1e-9)
    return ap.parse_args()


# ---------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [s.strip() for s in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); power.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


# ---------------------------------------------------------------- CPU legs (oracle = test infrastructure)
def cpu_pool_rate(bodies, seconds_target: float, workers: int):
    """Times the work-pool port (oracle/, kind 'port') on a bounded i-slice of the workload.
    Returns (interactions/s, sample description)."""
    from oracle.oracle import OracleSim
    o = OracleSim(bodies.copy())
    n = bodies.n
    # untimed: load the library, then ~2 s of threaded work so that the host cores reach their
    # steady clocks / placement (the first second of a burst runs at a fraction of the steady rate)
    t_w, rows_w = time.perf_counter(), min(n, max(128, 8 * workers))
    while time.perf_counter() - t_w < 2.0:
        o.time_slice(0, rows_w, workers)
    probe = min(n, max(128, 8 * workers))       # >= 100 rows: below that the reference runs a single slice
    while True:   # grow the probe until it is long enough to extrapolate from
        t0 = time.perf_counter()
        o.time_slice(0, probe, workers)
        dt = max(time.perf_counter() - t0, 1e-4)
        if dt >= 0.3 or probe >= n:
            break
        probe = min(n, probe * 4)
    rows = int(min(n, max(probe, probe * seconds_target / dt)))
    rows = max(workers, rows - rows % workers)
    t0 = time.perf_counter()
    o.time_slice(0, rows, workers)
    dt = time.perf_counter() - t0
    rate = rows * (n - 1) / dt
    return rate, f"{rows} i-bodies x {n} j-bodies (force sweep + collision sweep), {dt:.1f} s", rows, dt


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path.  The Go
    reference cannot be built here (no Go toolchain), so this times the C port of its
    goroutine work pool (oracle/, contiguous slices over all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nbodygo_b200 import clouds
    bodies = clouds.config(args.config, n=args.n or None)
    n = bodies.n
    workers = os.cpu_count() or 1
    per_step_s = 6.0
    for _ in range(max(args.warmup, 0)):
        cpu_pool_rate(bodies, 0.5, workers)
    tot_rows, tot_dt, sample = 0, 0.0, ""
    for _ in range(args.steps):
        _, sample, rows, dt = cpu_pool_rate(bodies, per_step_s, workers)
        tot_rows += rows
        tot_dt += dt
    value = tot_rows * (n - 1) / tot_dt
    line = {
        "impl": "reference", "metric": "fp64 body-pair interactions/s", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n * (n - 1) / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config}: {n}-body uniform sphere, elastic collisions", "n_bodies": n,
                   "note": "each step is a bounded i-slice of the workload; ms_per_step is the O(N^2) extrapolation "
                           "to a full cycle"},
        "steps_per_s": value / (n * (n - 1.0)),
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": workers, "kind": "port",
                         "sample": sample + f" per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(n: int, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_force launch from the committed
    `ncu --set full` capture of the same workload (profiles/r1_k_force_traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_k_force_traffic.json")) as f:
            for rec in json.load(f):
                if rec["n_bodies"] == n and rec["n_gpus"] == world:
                    return rec["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    return None


# ---------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nbodygo_b200 import capi, clouds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    bodies = clouds.config(args.config, n=args.n or None)   # same seed on every rank
    n = bodies.n
    sim = capi.Sim(n, device=local)
    sim.upload(bodies)
    if world > 1:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.comm_init(rank, world, uid[0])
    ts, R = args.time_scaling, 1.0
    # the roofline of K1 needs the CUDA events around it (nb_step_result.ms_force)
    opts = capi.STEP_DEFAULT | capi.STEP_PHASE_TIMINGS
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # fp64 roofline denominator measured on this device (MEASURED_PEAKS.json has no fp64 entry)
    peak_burst, _ = capi.measure_fp64_peak(local, 2048)
    barrier()
    for _ in range(args.warmup):
        sim.step(ts, R, opts)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = sim.launch_count()
    ms_steps, ms_force, pairs_seen, rounds = [], [], 0, 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        res = sim.step(ts, R, opts)          # CUDA events on the library's stream bracket the step
        ms_steps.append(res.ms_total)
        ms_force.append(res.ms_force)
        pairs_seen += res.n_pairs
        rounds = max(rounds, res.resolve_rounds)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = sim.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t_dev = max_over_ranks(sum(ms_steps) * 1e-3)
    t_force = max_over_ranks(float(np.mean(ms_force)) * 1e-3)
    interactions = float(n) * (n - 1.0)
    value = interactions * args.steps / t_dev
    peak_sust, _ = capi.measure_fp64_peak(local, 1 << 15)   # ~1 s back-to-back DFMA under the power cap

    # ---- e2e: host buffers through the C ABI every step -----------------------------------
    e2e = None
    if not args.no_e2e:
        names = ("x", "y", "z", "vx", "vy", "vz", "mass", "radius")
        pinned = {k: torch.empty(n, dtype=torch.float64).pin_memory().numpy() for k in names}
        for k in names:
            pinned[k][:] = getattr(bodies, k)
        beh = torch.empty(n, dtype=torch.uint8).pin_memory().numpy(); beh[:] = bodies.behavior
        flg = torch.empty(n, dtype=torch.uint8).pin_memory().numpy(); flg[:] = bodies.flags
        out = {k: torch.empty(n, dtype=torch.float64).pin_memory().numpy() for k in names[:6]}
        rxyz = torch.empty((n, 3), dtype=torch.float32).pin_memory().numpy()
        rex = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()

        def e2e_step():
            sim.upload_raw(n, *[pinned[k] for k in names], behavior=beh, flags=flg)
            sim.step(ts, R, opts)
            sim.download_into(**out)
            sim.render(rxyz, rex)
            for k in names[:6]:          # next cycle starts from the returned state, like the host loop:
                pinned[k], out[k] = out[k], pinned[k]   # both pinned — swap roles instead of a host memcpy

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": interactions * args.steps / t_e2e, "unit": "interactions/s",
               "h2d_bytes_per_step": int(n * (8 * 8 + 2)), "d2h_bytes_per_step": int(n * (6 * 8 + 13)),
               "ms_per_step": 1e3 * t_e2e / args.steps, "timer": "host perf_counter around synchronous C-ABI calls"}

    # ---- K1 with the uniform-mass pass disabled (N=1 only; reported beside the headline) ----
    # C4's bodies share one mass, so 31 of its 32 j-chunks run K1's uniform-mass instantiation
    # (DESIGN.md §3).  A collection with mixed masses runs the per-body-mass pass; its K1 time on this
    # same cloud is measured here, outside every timed region, so that both figures are on record.
    general = None
    if world == 1:
        try:
            os.environ["NB_UNIFORM_TILES"] = "0"
            try:
                sim_g = capi.Sim(n, device=local)      # the knob is read at nb_create
            finally:
                os.environ.pop("NB_UNIFORM_TILES", None)
            sim_g.upload(bodies)
            o_g = capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE | capi.STEP_PHASE_TIMINGS
            sim_g.step(ts, R, o_g)
            ms_g = min(sim_g.step(ts, R, o_g).ms_force for _ in range(2))
            sim_g.close()
            general = {"ms_per_launch": ms_g}
        except Exception as e:   # never let the side measurement cost the bench line
            general = {"error": str(e)}

    # ---- CPU baseline beside it (rank 0, N=1 only) -----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        rate, sample, _, _ = cpu_pool_rate(bodies, 15.0, workers)
        cpu = {"value": rate, "unit": "interactions/s", "cores": workers, "kind": "port", "sample": sample}

    if rank == 0:
        local_pairs = interactions / world
        achieved = FLOPS_PER_INTERACTION * local_pairs / t_force / 1e12
        k1_passes = ("uniform-mass instantiation on the j-chunks whose tiles each hold one mass, per-body-mass "
                     "instantiation on the rest")
        try:
            nu, nc, _, _ = capi.uniform_chunks(bodies)
            k1_passes += f" ({args.config}: {nu} of {nc} chunks uniform)"
        except Exception:
            pass
        if general and "ms_per_launch" in general:
            a_g = FLOPS_PER_INTERACTION * local_pairs / (general["ms_per_launch"] * 1e-3) / 1e12
            general.update(achieved=a_g, frac=a_g / peak_burst,
                           note="same launch with NB_UNIFORM_TILES=0: every chunk on the per-body-mass pass")
        line = {
            "metric": "fp64 body-pair interactions/s", "value": value, "unit": "interactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.config}: {n}-body uniform sphere, elastic collisions, one full cycle "
                                   "(force+detect+resolve+integrate)", "n_bodies": n, "parallelism": f"i-shard x{world}",
                       "l2": "flushed between timed steps (256 MiB write)", "seed": 3,
                       "collision_pairs_per_step": pairs_seen / max(args.steps, 1), "resolve_rounds": rounds},
            "steps_per_s": args.steps / t_dev,
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s",
                         "frac": achieved / peak_burst, "traffic": ncu_traffic(n, world),
                         "kernel": "k_force", "peak_source": "measured here: DFMA chain (nb_measure_fp64_peak), burst",
                         "peak_sustained": peak_sust, "frac_of_sustained": achieved / peak_sust,
                         "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
                         "flops_per_interaction": FLOPS_PER_INTERACTION,
                         "ms_per_launch": 1e3 * t_force,
                         "k1_passes": k1_passes,
                         "general_pass": general},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def relaunch_under_torchrun(args):
    """`python bench.py --gpus N` without a launcher: re-exec as one rank per GPU."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
    raise SystemExit(subprocess.call(cmd))


if __name__ == "__main__":
    a = parse()
    if a.gpus > 1 and "WORLD_SIZE" not in os.environ and a.impl == "ours":
        relaunch_under_torchrun(a)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
