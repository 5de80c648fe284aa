#!/usr/bin/env python3
"""Times K1 for alternative builds of nb_force.cu (-DNB_EXP_* knobs) on the GPU — the issue model of
tools/sass_model.py ranks them offline, this measures them.

  python tools/k1_hw_variants.py build          # here (no GPU): compiles nbodygo_b200/variants/*.so
  python tools/k1_hw_variants.py run [--n 256000]      # on the GPU box: one subprocess per variant

Each variant is timed on the C4 cloud with the uniform-mass pass (all chunks uniform) and with
NB_UNIFORM_TILES=0 (per-body-mass pass); best of 4 launches each."""
import json
import os

import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {          # the knobs that are still in the source (the rejected ones are in the history, see nb_force.cu)
    "prod": "",
    "nokreg": "-DNB_EXP_KREG=0",
    "kz_uni": "-DNB_EXP_KZ_UNI=1",
    "nokz_gen": "-DNB_EXP_KZ_GEN=0",
    "unr2": "-DNB_EXP_UNR4=2",
    "nstage2": "-DNB_EXP_NSTAGE=2",
}
FULL = ("NSTAGE",)     # knobs that live in nb_internal.cuh: rebuild every translation unit


def child(n):
    from nbodygo_b200 import capi, clouds
    b = clouds.config("C4", n=n)
    out = {}
    for uni in (1, 0):
        os.environ["NB_UNIFORM_TILES"] = str(uni)
        sim = capi.Sim(b.n)
        sim.upload(b)
        o = capi.STEP_COLLISIONS | capi.STEP_NO_INTEGRATE | capi.STEP_PHASE_TIMINGS
        sim.step(1e-9, 1.0, o)
        out["uni" if uni else "gen"] = min(sim.step(1e-9, 1.0, o).ms_force for _ in range(4))
        fx, fy, fz = sim.forces()
        out["chk_" + ("uni" if uni else "gen")] = [float(abs(fx).sum()), float(abs(fz).sum()), int(len(sim.pairs()))]
        sim.close()
    print(json.dumps(out))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "run"
    n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 256_000
    if mode == "child":
        return child(n)
    from nbodygo_b200 import _build
    if mode == "build":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sass_model
        for name, flags in VARIANTS.items():
            so = (_build.build_variant_full if any(k in flags for k in FULL) else _build.build_variant)(name, flags)
            unr = "2" if "UNR4=2" in flags else "1"
            rows = {r[0].split("ELi256ELi")[1][0]: (r[6], r[2], r[2] - r[4])
                    for r in sass_model.hot_loops(so, f"k_forceILi4ELi128ELi1ELi{unr}ELi256") if not r[3].get("SEL", 0)}
            print(f"{name:14s} {flags:60s} (model cycles, instr, non-FP64): uniform {rows.get('2')} general {rows.get('1')}", flush=True)
        return
    vdir = os.path.join(ROOT, "nbodygo_b200", "variants")
    base = None
    for name in VARIANTS:
        so = os.path.join(vdir, f"libnbody_b200_{name}.so")
        if not os.path.exists(so):
            continue
        env = dict(os.environ, NB_LIBRARY_PATH=so)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", "--n", str(n)], env=env,
                           capture_output=True, text=True)
        try:
            t = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:  # noqa: BLE001
            print(name, "failed:", r.stderr[-300:])
            continue
        base = base or t
        ok = all(abs(a - b) <= 1e-12 * abs(b) for k in ("chk_uni", "chk_gen") for a, b in zip(t[k][:2], base[k][:2])) \
            and t["chk_uni"][2] == base["chk_uni"][2] == t["chk_gen"][2]
        print(f"{name:14s} uniform {t['uni']:9.3f} ms ({100 * (t['uni'] / base['uni'] - 1):+5.2f} %)   "
              f"general {t['gen']:9.3f} ms ({100 * (t['gen'] / base['gen'] - 1):+5.2f} %)   "
              f"sum|F| and pair count agree with the first variant: {ok}", flush=True)


if __name__ == "__main__":
    main()
