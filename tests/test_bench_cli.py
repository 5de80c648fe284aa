"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys,
and the CUDA arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"}


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--bodies", "20000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "interactions/s" and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 0
    assert d["vs_baseline"] is None and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    # both arms describe the workload with the same `config` object (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config("C4", 20000, 1, 1e-9)


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--bodies", "2000"], capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0",
                        "--bodies", "1000"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_cuda_arm_json_assembly_with_a_stub_device(monkeypatch, capsys):
    """Runs bench.run_ours in-process with the device calls stubbed out (torch.cuda, capi.Sim and the
    peak probe): no number it prints means anything, but every line of the JSON assembly executes —
    a typo there would otherwise first show on the GPU box at round end."""
    import importlib
    import types

    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present: the real arm is exercised by the driver")
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    from nbodygo_b200 import capi

    class StubResult:
        ms_total, ms_force, n_pairs, resolve_rounds = 2.0, 1.5, 7, 2
        ms_exchange, ms_resolve, ms_integrate = 0.0, 0.2, 0.3

    class StubSim:
        created = 0

        def __init__(self, capacity, device=0, pair_capacity=0):
            StubSim.created += 1
            self.n, self._l = capacity, 0
            self.uniform = os.environ.get("NB_UNIFORM_TILES", "1")

        def upload(self, b): pass

        def upload_shard(self, n, first, count, *a, **k):
            assert (first, count) == (0, n) and len(a) == 8 and all(len(x) == n for x in a)

        def step(self, ts, R, opts=capi.STEP_DEFAULT):
            self._l += 5
            return StubResult()

        def download_range_into(self, first, count, **out): assert set(out) == {"x", "y", "z", "vx", "vy", "vz"}
        def render_range(self, first, count, xyz, ex): assert xyz.shape == (count, 3)
        def comm_mode(self): return capi.COMM_SINGLE
        def launch_count(self): return self._l
        def close(self): pass

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "device"}))
    monkeypatch.setattr(capi, "Sim", StubSim)
    monkeypatch.setattr(capi, "measure_fp64_peak", lambda dev, iters: (36.7, 1.0))
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="ours", config="C4", n=20000,
                                 no_cpu_baseline=True, no_e2e=False, time_scaling=1e-9, no_extra_configs=False,
                                 no_parity_check=False)
    bench.run_ours(args)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert (REQUIRED - {"impl"}) <= set(d)
    assert {"roofline", "gpu_launches", "clocks", "steps_per_s", "exchange", "ms_exchange", "parity_check", "configs",
            "run_stats"} <= set(d)
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_source", "general_pass", "k1_passes"} <= set(rf)
    assert rf["general_pass"]["ms_per_launch"] == 1.5 and "frac" in rf["general_pass"]
    assert "of" in rf["k1_passes"] and "chunks uniform" in rf["k1_passes"]
    assert d["e2e"]["h2d_bytes_per_step"] == 20000 * 66 and d["e2e"]["d2h_bytes_per_step"] == 20000 * 61
    assert d["gpu_launches"] == 10 and d["n_gpus"] == 1 and d["cpu_baseline"] is None
    assert d["exchange"] == "single" and d["parity_check"] is None and d["configs"] == []   # --bodies: headline only
    assert StubSim.created == 2 and "NB_UNIFORM_TILES" not in os.environ   # the side measurement cleaned up
    assert np.isclose(d["value"], 20000 * 19999 * 2 / (2 * 2.0e-3))
    # both arms describe the workload with the same `config` object
    assert d["config"] == bench.workload_config("C4", 20000, 1, 1e-9)

    # the default run also carries the other BASELINE configs (C2, C3, the C5 sweep) in `configs`
    args.n, args.no_e2e = 0, True
    bench.run_ours(args)
    d = json.loads([l for l in capsys.readouterr().out.splitlines() if l.strip()][0])
    names = [(c["config"], c["n_bodies"]) for c in d["configs"]]
    assert ("C2", 10_000) in names and ("C3", 100_000) in names and ("C5 sweep: C4", 256_000) in names
    assert all("error" not in c and c["value"] > 0 and 0 < c["roofline_frac"] for c in d["configs"])
    assert d["roofline"]["traffic"] is None or "profiles/" in d["roofline"]["traffic_source"]
