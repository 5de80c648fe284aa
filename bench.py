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
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the C2/C3/C5 lines of the `configs` array")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the untimed sharded-vs-1-GPU self-check")
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


WORKLOADS = {
    "C1": "reference sphere-cloud sim geometry (Sim3: sun + two clusters), elastic collisions",
    "C2": "uniform sphere, per-body masses, collisions off (pure all-pairs force + integrate)",
    "C3": "cube cloud, elastic collisions",
    "C3dense": "cube cloud with 4x radii (collision-dense), elastic collisions",
    "C4": "uniform sphere, elastic collisions",
}
SEEDS = {"C1": 11, "C2": 1, "C3": 2, "C3dense": 2, "C4": 3}


def workload_config(name: str, n: int, world: int, ts: float):
    """The `config` object of the JSON line — identical in both arms (ours / --impl reference)."""
    return {"workload": f"{name}: {n}-body {WORKLOADS.get(name, name)}, one full cycle "
                        "(force+detect+resolve+integrate)",
            "n_bodies": n, "parallelism": f"i-shard x{world}", "seed": SEEDS.get(name),
            "time_scaling": ts, "R": 1.0,
            "l2": "GPU arm: flushed between timed steps (256 MiB write)"}


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
    per_step_s = 5.0
    for _ in range(max(args.warmup, 0)):
        cpu_pool_rate(bodies, 0.5, workers)
    tot_rows, tot_dt, sample = 0, 0.0, ""
    for _ in range(args.steps):
        _, sample, rows, dt = cpu_pool_rate(bodies, per_step_s, workers)
        tot_rows += rows
        tot_dt += dt
    # Body.Update over the whole array (O(n), once per cycle) timed for real and added to the extrapolated
    # cycle; ProcessMods handles ~1e3 events per cycle at C4 (microseconds) and is not in the sample
    from oracle.oracle import OracleSim
    ou = OracleSim(bodies.copy())
    t0 = time.perf_counter()
    ou.update(args.time_scaling, 1.0)
    t_update = time.perf_counter() - t0
    t_cycle = n * (n - 1.0) * tot_dt / (tot_rows * (n - 1.0)) + t_update
    value = n * (n - 1.0) / t_cycle
    line = {
        "impl": "reference", "metric": "fp64 body-pair interactions/s", "value": value, "unit": "interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_cycle, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.config, n, args.gpus, args.time_scaling),
        "note": "the Go reference cannot be built here (no Go toolchain): this is the C port of its goroutine work "
                "pool (oracle/, contiguous slices over all host threads).  Each step is a bounded i-slice of the "
                "workload (Body.Compute: force sweep + collision sweep); ms_per_step is the O(N^2) extrapolation "
                f"to a full cycle plus Body.Update over all bodies measured for real ({1e3 * t_update:.1f} ms)",
        "steps_per_s": 1.0 / t_cycle,
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": workers, "kind": "port",
                         "sample": sample + f" per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(n: int, world: int):
    """(dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch, where that number comes from): read from
    the committed `ncu --set full` capture of the same workload and kernel instantiation
    (profiles/r2_k_force_traffic.json) — bench.py itself never runs under a profiler."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_k_force_traffic.json")) as f:
            for rec in json.load(f):
                if rec["n_bodies"] == n and rec["n_gpus"] == world:
                    return rec["dram_bytes_per_launch"], f"profiles/{rec['report']} ({rec['kernel']}), not measured in this run"
    except (OSError, ValueError, KeyError):
        pass
    return None, None


def digest(*arrays) -> str:
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


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

    def new_sim(bodies, shard_upload=False):
        """One handle per rank holding `bodies`, joined into a communicator when world > 1."""
        sim = capi.Sim(bodies.n, device=local)
        if world == 1:
            sim.upload(bodies)
            return sim
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        if not shard_upload:
            sim.upload(bodies)
        sim.comm_init(rank, world, uid[0])
        return sim

    ts, R = args.time_scaling, 1.0
    # the roofline of K1 needs the CUDA events around it (nb_step_result.ms_force)
    opts = capi.STEP_DEFAULT | capi.STEP_PHASE_TIMINGS
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    interactions_of = lambda n_: float(n_) * (n_ - 1.0)

    def timed_cycles(sim, steps, warmup, opts=opts):
        """`warmup` untimed cycles, then `steps` cycles timed by the CUDA events the library records on its
        own stream around nb_step (L2 flushed before each); returns max-over-ranks seconds and per-phase means."""
        for _ in range(warmup):
            sim.step(ts, R, opts)
        barrier()
        acc = dict(total=[], force=[], exchange=[], resolve=[], integrate=[], pairs=0, rounds=0)
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            res = sim.step(ts, R, opts)
            acc["total"].append(res.ms_total); acc["force"].append(res.ms_force)
            acc["exchange"].append(res.ms_exchange); acc["resolve"].append(res.ms_resolve)
            acc["integrate"].append(res.ms_integrate)
            acc["pairs"] += res.n_pairs
            acc["rounds"] = max(acc["rounds"], res.resolve_rounds)
        barrier()
        if steps:
            acc["t_dev"] = max_over_ranks(sum(acc["total"]) * 1e-3)
            acc["t_force"] = max_over_ranks(float(np.mean(acc["force"])) * 1e-3)
        return acc


    bodies = clouds.config(args.config, n=args.n or None)   # same seed on every rank
    n = bodies.n
    sim = new_sim(bodies)
    interactions = interactions_of(n)

    # fp64 roofline denominator measured on this device (MEASURED_PEAKS.json has no fp64 entry)
    peak_burst, _ = capi.measure_fp64_peak(local, 2048)
    barrier()

    # ---- parity self-check (world > 1, untimed): the first sharded cycle against one GPU ----------------
    # Every rank digests the forces of its shard, the gathered pair list and the whole state after the cycle;
    # rank 0 steps a single-GPU handle from the same initial state and compares (bit equality through SHA-256).
    # The driver's GPU test box has one GPU, so this is where the sharded cycle is checked at every N.
    parity = None
    warm_done = 0
    if world > 1 and not args.no_parity_check:
        sim.step(ts, R, opts)
        warm_done = 1
        i0, i1 = sim.shard_range()
        fx, fy, fz = sim.forces()
        st = sim.download()
        mine = {"i0": i0, "i1": i1, "force": digest(fx[i0:i1], fy[i0:i1], fz[i0:i1]), "pairs": digest(sim.pairs()),
                "state": digest(st.x, st.y, st.z, st.vx, st.vy, st.vz, st.mass, st.rest, st.flags, st.behavior)}
        del fx, fy, fz, st
        got = [None] * world
        dist.gather_object(mine, got if rank == 0 else None, dst=0)
        if rank == 0:
            ref = capi.Sim(n, device=local)
            ref.upload(bodies)
            ref.step(ts, R, opts)
            rfx, rfy, rfz = ref.forces()
            rst = ref.download()
            r_pairs = digest(ref.pairs())
            r_state = digest(rst.x, rst.y, rst.z, rst.vx, rst.vy, rst.vz, rst.mass, rst.rest, rst.flags, rst.behavior)
            parity = {
                "force_bits_equal": all(g["force"] == digest(rfx[g["i0"]:g["i1"]], rfy[g["i0"]:g["i1"]],
                                                             rfz[g["i0"]:g["i1"]]) for g in got),
                "pairs_equal": all(g["pairs"] == r_pairs for g in got),
                "state_bits_equal": all(g["state"] == r_state for g in got),
                "ranks_checked": world, "n_pairs": int(len(ref.pairs())),
                "what": "cycle 1 from the uploaded state: each rank's shard forces, gathered pair list and full state "
                        "(x..vz, mass, rest, flags, behavior) vs a single-GPU handle on rank 0, SHA-256 of the raw bytes",
            }
            ref.close()
            del rfx, rfy, rfz, rst
        barrier()

    timed_cycles(sim, 0, max(args.warmup - warm_done, 0))      # the remaining warm-up cycles
    sampler = ClockSampler(local) if rank == 0 else None       # clocks / throttle reasons DURING the timed region
    launches0 = sim.launch_count()
    t_wall0 = time.perf_counter()
    acc = timed_cycles(sim, args.steps, 0)
    t_wall = time.perf_counter() - t_wall0
    launches = sim.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t_dev, t_force = acc["t_dev"], acc["t_force"]
    value = interactions * args.steps / t_dev
    peak_sust, _ = capi.measure_fp64_peak(local, 1 << 15)   # ~1 s back-to-back DFMA under the power cap
    mode = capi.COMM_MODE_NAMES.get(sim.comm_mode(), "?")

    # ---- e2e: host buffers through the C ABI every step ---------------------------------------------------
    # Every rank moves only its own i-range over its host link (nb_upload_shard / nb_download_*_range): the other
    # ranks' slices arrive over NVLink inside the upload, and after the cycle every rank returns its own slice.
    e2e = None
    if not args.no_e2e:
        i0, i1, _, _ = capi.plan(n, rank, world)
        cnt = i1 - i0
        names = ("x", "y", "z", "vx", "vy", "vz", "mass", "radius")
        pinned = {k: torch.empty(cnt, dtype=torch.float64).pin_memory().numpy() for k in names}
        for k in names:
            pinned[k][:] = getattr(bodies, k)[i0:i1]
        beh = torch.empty(cnt, dtype=torch.uint8).pin_memory().numpy(); beh[:] = bodies.behavior[i0:i1]
        flg = torch.empty(cnt, dtype=torch.uint8).pin_memory().numpy(); flg[:] = bodies.flags[i0:i1]
        out = {k: torch.empty(cnt, dtype=torch.float64).pin_memory().numpy() for k in names[:6]}
        rxyz = torch.empty((cnt, 3), dtype=torch.float32).pin_memory().numpy()
        rex = torch.empty(cnt, dtype=torch.uint8).pin_memory().numpy()

        def e2e_step():
            sim.upload_shard(n, i0, cnt, *[pinned[k] for k in names], behavior=beh, flags=flg)
            sim.step(ts, R, opts)
            sim.download_range_into(i0, cnt, **out)
            sim.render_range(i0, cnt, rxyz, rex)
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
               "ms_per_step": 1e3 * t_e2e / args.steps, "timer": "host perf_counter around synchronous C-ABI calls",
               "path": ("nb_upload_shard -> nb_step -> nb_download_state_range + nb_download_render_range; the byte "
                        "counts are totals over all ranks (each rank moves its own n/N slice over its host link)")}
    sim.close()

    # ---- K1 with the uniform-mass pass disabled (N=1 only; reported beside the headline) ----
    # C4's bodies share one mass, so its j-chunks run K1's uniform-mass instantiation.  A
    # collection with mixed masses runs the per-body-mass pass; its K1 time on this same cloud is measured
    # here, outside every timed region, so that both figures are on record.
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

    # ---- the other BASELINE configs on the same code, same protocol (untimed for the headline) -------------
    extra = []
    if not args.no_extra_configs and not args.n:
        plan = [("C2", None, 30), ("C3", None, 10), ("C3dense", None, 5)]
        sweep = [1_000, 4_000, 16_000, 64_000, 256_000] + ([4_000_000] if world >= 4 else [])
        plan += [("C4", m, 30 if m <= 16_000 else (8 if m <= 256_000 else 2)) for m in sweep]
        for name, m, k in plan:
            try:
                b2 = clouds.config(name, n=m)
                s2 = new_sim(b2)
                # C2 is "collisions off": no detect / resolve in its cycle (the overlap mask of the force stays).
                # The cycle is timed WITHOUT the per-phase events (each is a node between two kernels and costs
                # a small cycle 2-3 us; a single-GPU cycle then replays its CUDA graph, as in production); K1's
                # time for the roofline comes from three more cycles with them.
                base_opts = 0 if name == "C2" else capi.STEP_COLLISIONS
                a2 = timed_cycles(s2, k, 3, base_opts)
                slow = a2["t_dev"] / k > 0.5                      # 4 M bodies: one more cycle is enough
                a2["t_force"] = timed_cycles(s2, 1 if slow else 3, 0 if slow else 1,
                                             base_opts | capi.STEP_PHASE_TIMINGS)["t_force"]
                s2.close()
                it = interactions_of(b2.n)
                ach = FLOPS_PER_INTERACTION * (it / world) / a2["t_force"] / 1e12
                extra.append({"config": ("C5 sweep: " if m else "") + name, "n_bodies": b2.n, "n_gpus": world, "steps": k,
                              "warmup": 3, "value": it * k / a2["t_dev"], "unit": "interactions/s",
                              "ms_per_step": 1e3 * a2["t_dev"] / k, "steps_per_s": k / a2["t_dev"],
                              "ms_force": 1e3 * a2["t_force"], "roofline_frac": ach / max(peak_burst, peak_sust),
                              "collision_pairs_per_step": a2["pairs"] / k, "resolve_rounds": a2["rounds"]})
            except Exception as e:
                extra.append({"config": name, "n_bodies": m, "error": str(e)})
            barrier()

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
            general.update(achieved=a_g, frac=a_g / max(peak_burst, peak_sust),
                           note="same launch with NB_UNIFORM_TILES=0: every chunk on the per-body-mass pass")
        # roofline denominator: the larger of the two DFMA-chain measurements of this run (a 4 ms burst before the
        # timed region and ~70 ms back to back after it).  The short one scatters by +-0.5 % from run to run; taking
        # the maximum never lets a low probe flatter the kernel.
        peak = max(peak_burst, peak_sust)
        traffic, traffic_source = ncu_traffic(n, world)
        cfg = workload_config(args.config, n, world, ts)
        line = {
            "metric": "fp64 body-pair interactions/s", "value": value, "unit": "interactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "run_stats": {"collision_pairs_per_step": acc["pairs"] / max(args.steps, 1), "resolve_rounds": acc["rounds"],
                          "ms_force": float(np.mean(acc["force"])), "ms_exchange": float(np.mean(acc["exchange"])),
                          "ms_resolve": float(np.mean(acc["resolve"])), "ms_integrate": float(np.mean(acc["integrate"]))},
            "exchange": mode, "ms_exchange": float(np.mean(acc["exchange"])),
            "parity_check": parity,
            "steps_per_s": args.steps / t_dev,
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                         "kernel": "k_force",
                         "peak_source": "measured here: DFMA chain (nb_measure_fp64_peak), max(burst, sustained)",
                         "peak_burst": peak_burst, "frac_of_burst": achieved / peak_burst,
                         "peak_sustained": peak_sust, "frac_of_sustained": achieved / peak_sust,
                         "peak_nominal": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
                         "flops_per_interaction": FLOPS_PER_INTERACTION,
                         "ms_per_launch": 1e3 * t_force,
                         "k1_passes": k1_passes,
                         "general_pass": general},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "configs": extra,
        }
        print(json.dumps(line), flush=True)
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
