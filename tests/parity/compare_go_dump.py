#!/usr/bin/env python3
"""Pins the CPU oracle (and, with --gpu, the CUDA path) against a dump of the REAL reference.

The dump is written on a box with a Go toolchain by integration/cmd/sim/parity_dump_test.go (the unmodified
Body.Compute → ProcessMods → Body.Update → Cycle of aceeric/nbodygo, single worker) from a CSV that
tools/write_inputs.py wrote.  This script replays the same CSV through oracle/ and reports, per cycle:

  events   the raw reference stream (self pairs and dead-j events included), order included   -> must be equal
  forces   Body.fx,fy,fz after Compute                                                        -> must be bit-equal
           (Go gc on amd64 does not fuse multiply-add; on arm64 it does and the last bits move)
  state    X..Vz, Mass, Exists after Update, for both transcendental backends of the oracle:
           ORC_MATH_GO (Go's math restated, oracle/gomath.c) should be bit-equal; glibc within a few ulp
           except near head-on collisions

  python tests/parity/compare_go_dump.py --csv c3_2000.csv --dump go_dump.txt [--gpu]
  python tests/parity/compare_go_dump.py --self-test      # no Go needed: the oracle writes the dump itself

This file lives under tests/ because it drives oracle/ (test infrastructure).  Exit code 0 = pinned.
"""
from __future__ import annotations

import argparse
import os
import struct
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from nbodygo_b200 import clouds  # noqa: E402
from oracle import oracle as orc  # noqa: E402

STATE = ("x", "y", "z", "vx", "vy", "vz", "mass")


def hx(v: float) -> str:
    return struct.pack(">d", float(v)).hex()


def unhx(s: str) -> float:
    return struct.unpack(">d", bytes.fromhex(s))[0]


def ulps(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units of the last place between two float64 arrays (NaN == NaN counts as 0)."""
    ia = a.view(np.int64).copy()
    ib = b.view(np.int64).copy()
    ia[ia < 0] = np.int64(-(2 ** 63)) - ia[ia < 0]
    ib[ib < 0] = np.int64(-(2 ** 63)) - ib[ib < 0]
    with np.errstate(over="ignore"):
        d = np.abs(ia - ib).astype(np.float64)   # exact for any pair of like-signed values
    d[np.isnan(a) & np.isnan(b)] = 0
    return d


# ---------------------------------------------------------------- the dump format
def parse_dump(path):
    """-> header dict, list of cycles: dict(n, events [(kind,a,b)], forces [n,3], state {f: [n]}, exists [n], n_after)"""
    header, cycles, cur = None, [], None
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "H":
                header = dict(n=int(t[1]), ts=unhx(t[2]), R=unhx(t[3]), cycles=int(t[4]))
            elif t[0] == "C":
                n = int(t[2])
                cur = dict(n=n, events=[], forces=np.zeros((n, 3)), state={k: np.zeros(n) for k in STATE},
                           exists=np.zeros(n, dtype=bool), n_after=None)
                cycles.append(cur)
            elif t[0] == "E":
                cur["events"].append((int(t[1]), int(t[2]), int(t[3])))
            elif t[0] == "F":
                cur["forces"][int(t[1])] = [unhx(v) for v in t[2:5]]
            elif t[0] == "S":
                i = int(t[1])
                for k, v in zip(STATE, t[2:9]):
                    cur["state"][k][i] = unhx(v)
                cur["exists"][i] = t[9] == "1"
            elif t[0] == "N":
                cur["n_after"] = int(t[1])
    return header, cycles


def write_dump_from_oracle(path, bodies, ts, R, cycles, math_backend):
    """The dump parity_dump_test.go writes, produced by the oracle instead (self-test / format reference)."""
    prev = orc.set_math(math_backend)
    try:
        o = orc.OracleSim(bodies.copy())
        with open(path, "w") as f:
            f.write(f"H {o.b.n} {hx(ts)} {hx(R)} {cycles}\n")
            for c in range(cycles):
                n = o.b.n
                f.write(f"C {c} {n}\n")
                o.compute(opts=orc.OPT_SELF_PAIRS | orc.OPT_DEAD_J)
                for e in o.events[::-1]:     # PushFront + Front→Next: reverse arrival order
                    f.write(f"E {int(e['kind'])} {int(e['a'])} {int(e['b'])}\n")
                for i in range(n):
                    f.write(f"F {i} {hx(o.fx[i])} {hx(o.fy[i])} {hx(o.fz[i])}\n")
                o.process_mods()
                o.update(ts, R)
                for i in range(n):
                    f.write("S %d %s %d\n" % (i, " ".join(hx(getattr(o.b, k)[i]) for k in STATE), int(o.b.exists[i])))
                o.cycle_compact()
                f.write(f"N {o.b.n}\n")
    finally:
        orc.set_math(prev)


# ---------------------------------------------------------------- the comparison
def replay(bodies, header, cycles, backend, log):
    """Replays the dump's cycles through the oracle with one transcendental backend; returns a summary dict."""
    prev = orc.set_math(backend)
    name = "go-math" if backend == orc.MATH_GO else "glibc"
    out = dict(events_equal=True, force_bits_equal=True, state_bits_equal=True, exists_equal=True,
               max_state_ulp=0.0, max_state_rel=0.0, n_equal=True)
    try:
        o = orc.OracleSim(bodies.copy())
        for c, cyc in enumerate(cycles):
            if cyc["n"] != o.b.n:
                out["n_equal"] = False
                log(f"[{name}] cycle {c}: body count {o.b.n} vs reference {cyc['n']} — stopping")
                break
            o.compute(opts=orc.OPT_SELF_PAIRS | orc.OPT_DEAD_J)
            mine = [(int(e["kind"]), int(e["a"]), int(e["b"])) for e in o.events[::-1]]
            if mine != cyc["events"]:
                out["events_equal"] = False
                same_set = sorted(mine) == sorted(cyc["events"])
                log(f"[{name}] cycle {c}: event stream differs ({len(mine)} vs {len(cyc['events'])} events; "
                    f"{'same set, different order' if same_set else 'different sets'})")
            f = np.stack([o.fx, o.fy, o.fz], axis=1)
            fb = f.view(np.uint64) == cyc["forces"].view(np.uint64)
            if not fb.all():
                out["force_bits_equal"] = False
                log(f"[{name}] cycle {c}: {int((~fb).sum())} force words differ, max {ulps(f, cyc['forces']).max():.0f} ulp")
            o.process_mods()
            o.update(header["ts"], header["R"])
            if not np.array_equal(o.b.exists, cyc["exists"]):
                out["exists_equal"] = False
                log(f"[{name}] cycle {c}: Exists differs for {int((o.b.exists != cyc['exists']).sum())} bodies")
            for k in STATE:
                a, r = getattr(o.b, k), cyc["state"][k]
                if not np.array_equal(a.view(np.uint64), r.view(np.uint64)):
                    out["state_bits_equal"] = False
                    u = ulps(a, r)
                    scale = max(np.nanmax(np.abs(r)), 1e-300)
                    with np.errstate(invalid="ignore"):
                        rel = np.nanmax(np.abs(a - r)) / scale
                    out["max_state_ulp"] = max(out["max_state_ulp"], float(u.max()))
                    out["max_state_rel"] = max(out["max_state_rel"], float(rel))
                    log(f"[{name}] cycle {c}: {k}: {int((u > 0).sum())} values differ, max {u.max():.0f} ulp, "
                        f"{rel:.2e} of the largest |{k}|")
            o.cycle_compact()
            if cyc["n_after"] is not None and cyc["n_after"] != o.b.n:
                out["n_equal"] = False
                log(f"[{name}] cycle {c}: count after Cycle {o.b.n} vs reference {cyc['n_after']}")
    finally:
        orc.set_math(prev)
    return out


def compare_gpu(bodies, header, cycles, log):
    """The CUDA path against the reference dump: canonical pair set bit-exact, forces within 1e-12 normwise."""
    from nbodygo_b200 import capi
    sim = capi.Sim(bodies.n + 16)
    sim.upload(bodies)
    o = orc.OracleSim(bodies.copy())
    ok = True
    for c, cyc in enumerate(cycles):
        if sim.count() != cyc["n"]:
            log(f"[gpu] cycle {c}: body count {sim.count()} vs reference {cyc['n']} — stopping")
            return False
        alive = (o.b.flags & 1) != 0
        ref_pairs = sorted((a, b) for k, a, b in cyc["events"] if k == 0 and a != b and alive[b])
        _, _, _, fn = o.compute_exact()
        sim.step(header["ts"], header["R"])
        got = [tuple(p) for p in sim.pairs()]
        if got != ref_pairs:
            ok = False
            log(f"[gpu] cycle {c}: collision pair set differs ({len(got)} vs {len(ref_pairs)})")
        f = np.stack(sim.forces(), axis=1)
        err = np.abs(f - cyc["forces"]).max(axis=1)
        worst = float(np.max(err[fn > 0] / fn[fn > 0])) if (fn > 0).any() else 0.0
        log(f"[gpu] cycle {c}: max normwise force deviation from the reference {worst:.2e} (bar 1e-12)")
        ok = ok and worst <= 1e-12
        st = sim.download()
        for k in ("x", "vx"):
            r = cyc["state"][k]
            with np.errstate(invalid="ignore"):
                rel = np.nanmax(np.abs(getattr(st, k) - r)) / max(np.nanmax(np.abs(r)), 1e-300)
            log(f"[gpu] cycle {c}: {k} within {rel:.2e} of the largest |{k}|")
            ok = ok and rel <= 1e-9
        # keep the oracle (used for norms and liveness) in step with the reference
        o.compute(opts=orc.OPT_SELF_PAIRS | orc.OPT_DEAD_J)
        o.process_mods()
        o.update(header["ts"], header["R"])
        o.cycle_compact()
        sim.compact()
    sim.close()
    return ok


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--csv")
    ap.add_argument("--dump")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--self-test", action="store_true")
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args(argv)
    log = (lambda *_: None) if a.quiet else print
    if a.self_test:
        tmp = tempfile.mkdtemp()
        b = clouds.config("C1", n=301)
        from nbodygo_b200.bodies import F_EXISTS
        b.flags[5] &= ~np.uint8(F_EXISTS)      # a dead body in the array: exercises the dead-j events
        a.csv = os.path.join(tmp, "c1_301.csv")
        a.dump = os.path.join(tmp, "dump.txt")
        clouds.write_csv(a.csv, b)
        bodies = clouds.read_csv(a.csv)
        bodies.flags[5] &= ~np.uint8(F_EXISTS)
        write_dump_from_oracle(a.dump, bodies, 1e-9, 1.0, 3, orc.MATH_GO)
    else:
        if not a.csv or not a.dump:
            ap.error("--csv and --dump are required (or --self-test)")
        bodies = clouds.read_csv(a.csv)
    header, cycles = parse_dump(a.dump)
    assert header and header["n"] == bodies.n, "the dump was not taken from this CSV"
    res_go = replay(bodies, header, cycles, orc.MATH_GO, log)
    res_libm = replay(bodies, header, cycles, orc.MATH_LIBM, log)
    pinned = all(res_go[k] for k in ("events_equal", "force_bits_equal", "exists_equal", "n_equal"))
    print(f"oracle (Go math backend): events_equal={res_go['events_equal']} force_bits_equal={res_go['force_bits_equal']} "
          f"state_bits_equal={res_go['state_bits_equal']} (max {res_go['max_state_ulp']:.0f} ulp, "
          f"{res_go['max_state_rel']:.2e} rel)")
    print(f"oracle (glibc backend):   events_equal={res_libm['events_equal']} force_bits_equal={res_libm['force_bits_equal']} "
          f"state_bits_equal={res_libm['state_bits_equal']} (max {res_libm['max_state_ulp']:.0f} ulp, "
          f"{res_libm['max_state_rel']:.2e} rel)")
    if a.gpu:
        pinned = compare_gpu(bodies, header, cycles, log) and pinned
    print("PINNED" if pinned else "NOT PINNED")
    return 0 if pinned else 1


if __name__ == "__main__":
    raise SystemExit(main())
