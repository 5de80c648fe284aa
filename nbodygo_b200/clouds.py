"""Seeded input clouds and the CSV identical-input channel.

The reference's built-in generators are unseeded (``rand.Seed(time.Now()...)`` in
cmd/util/vectorutil.go:34 and cmd/sim/simgen.go:114,183,258), so identical inputs
can only reach the Go server through its CSV loader (cmd/sim/fromcsv.go:15-47).
These generators restate the *geometry* of the reference sims with a recorded
seed, and ``write_csv`` emits the 13-column format FromCsv parses (``%.17g``).
Configs C1..C4 are the synthetic inputs of SURVEY.md §8(d).
"""
from __future__ import annotations

import numpy as np

from .bodies import (BEHAVIOR_NAMES, ELASTIC, F_EXISTS, F_PINNED, F_SUN, NONE, SUBSUME, BodyArrays,
                     parse_collision_behavior)


def _in_sphere(rng, n, radius, center=(0.0, 0.0, 0.0)):
    """Uniform points in a ball (the rejection rule of util.GetVectorEven, vectorised)."""
    out = np.empty((0, 3))
    while len(out) < n:
        p = rng.uniform(-1.0, 1.0, size=(int((n - len(out)) * 2.2) + 16, 3))
        p = p[(p * p).sum(axis=1) <= 1.0]
        out = np.concatenate([out, p])
    return out[:n] * radius + np.asarray(center)


def uniform_sphere(n, radius, body_radius, mass, vmax=0.0, behavior=ELASTIC, seed=0, mass_hi=None):
    rng = np.random.default_rng(seed)
    p = _in_sphere(rng, n, radius)
    v = rng.uniform(-vmax, vmax, size=(n, 3)) if vmax else np.zeros((n, 3))
    m = np.full(n, mass) if mass_hi is None else rng.uniform(mass, mass_hi, n)
    return BodyArrays.from_fields(p[:, 0], p[:, 1], p[:, 2], v[:, 0], v[:, 1], v[:, 2], m,
                                  np.full(n, body_radius), behavior=behavior)


def uniform_cube(n, side, body_radius, mass, vmax=0.0, behavior=ELASTIC, seed=0):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-side / 2, side / 2, size=(n, 3))
    v = rng.uniform(-vmax, vmax, size=(n, 3)) if vmax else np.zeros((n, 3))
    return BodyArrays.from_fields(p[:, 0], p[:, 1], p[:, 2], v[:, 0], v[:, 1], v[:, 2], np.full(n, mass),
                                  np.full(n, body_radius), behavior=behavior)


def sim3_like(n, seed=0, cluster_radius=50.0, mass=90e25, behavior=ELASTIC):
    """Geometry of generator.Sim3 (cmd/sim/simgen.go:238-254): a far sun + two clusters."""
    rng = np.random.default_rng(seed)
    half = (n - 1) // 2
    parts = [BodyArrays.from_fields([1e5], [1e5], [1e5], [-3.0], [-3.0], [-5.0], [1.0], [500.0],
                                    behavior=SUBSUME, flags=F_EXISTS | F_SUN | F_PINNED)]
    for j in (-1.0, 1.0):
        p = _in_sphere(rng, half, cluster_radius, (j * 70, j * 70, j * 70))
        parts.append(BodyArrays.from_fields(p[:, 0], p[:, 1], p[:, 2], np.full(half, j * 121185000.0),
                                            np.full(half, j * 121185000.0), np.full(half, j * -121185000.0),
                                            np.full(half, mass), np.full(half, 5.0), behavior=behavior))
    b = parts[0]
    for q in parts[1:]:
        b.append(q)
    b.id = np.arange(b.n, dtype=np.int64)
    return b


# ---- the named configs of BASELINE.json / SURVEY §8(d) ------------------------
def config(name: str, n: int | None = None) -> BodyArrays:
    if name == "C1":   # reference sphere-cloud sim, ~1,000 bodies, elastic
        return sim3_like(n or 1001, seed=11)
    if name == "C2":   # 10,000-body uniform sphere, collisions off
        return uniform_sphere(n or 10_000, 1000.0, 1e-3, 1e24, behavior=NONE, seed=1, mass_hi=1e25)
    if name == "C3":   # 100,000-body cube cloud with elastic collisions
        return uniform_cube(n or 100_000, 2000.0, 1.684, 1e24, vmax=1e8, seed=2)
    if name == "C3dense":
        return uniform_cube(n or 100_000, 2000.0, 4 * 1.684, 1e24, vmax=1e8, seed=2)
    if name == "C4":   # 1,000,000-body uniform sphere with elastic collisions
        nn = n or 1_000_000
        # keep the expected number of overlapping pairs at ~1e-3 * n when n is scaled down
        r = 3.150 * (1_000_000 / nn) ** (1.0 / 3.0)
        return uniform_sphere(nn, 5000.0, r, 1e24, vmax=1e8, seed=3)
    raise ValueError(f"unknown config {name}")


# ---- CSV channel (cmd/sim/fromcsv.go:15-47) ------------------------------------
def write_csv(path: str, b: BodyArrays) -> None:
    with open(path, "w") as f:
        f.write("# x,y,z,vx,vy,vz,mass,radius,is_sun,collision_behavior,color,frag_factor,frag_step\n")
        for i in range(b.n):
            nums = ",".join("%.17g" % getattr(b, k)[i] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius"))
            sun = "true" if b.flags[i] & F_SUN else "false"
            f.write(f"{nums},{sun},{BEHAVIOR_NAMES[b.behavior[i]]},white,"
                    f"{'%.17g' % b.frag_factor[i]},{'%.17g' % b.frag_step[i]}\n")


def read_csv(path: str, body_count: int = 1 << 62, default_behavior: int = ELASTIC) -> BodyArrays:
    """FromCsv semantics: '#' comments, optional trailing fields, bad rows skipped."""
    rows = []
    with open(path) as f:
        for line in f:
            if len(rows) >= body_count:
                break
            if line.startswith("#") or not line.strip():
                continue
            fld = [s.strip() for s in line.rstrip("\n").split(",")]
            try:
                vals = [float(fld[k]) for k in range(8)]
                sun = False
                if len(fld) >= 9:
                    if fld[8].lower() in ("1", "t", "true"):
                        sun = True
                    elif fld[8].lower() in ("0", "f", "false"):
                        sun = False
                    else:
                        raise ValueError("bad bool")  # strconv.ParseBool error ⇒ row skipped
                beh = parse_collision_behavior(fld[9]) if len(fld) >= 10 else default_behavior
                ff = float(fld[11]) if len(fld) >= 12 else 0.0
                fs = float(fld[12]) if len(fld) >= 13 else 0.0
            except (ValueError, IndexError):
                continue
            rows.append((vals, sun, beh, ff, fs))
    b = BodyArrays(len(rows))
    for i, (vals, sun, beh, ff, fs) in enumerate(rows):
        for k, name in enumerate(("x", "y", "z", "vx", "vy", "vz", "mass", "radius")):
            getattr(b, name)[i] = vals[k]
        b.behavior[i] = beh
        b.flags[i] = F_EXISTS | (F_SUN if sun else 0)
        b.frag_factor[i], b.frag_step[i] = ff, fs
    return b
