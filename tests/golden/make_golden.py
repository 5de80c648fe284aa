#!/usr/bin/env python3
"""Generates tests/golden/golden.json — an independent restatement of the hot path.

The reference (Go) cannot run in this image, so these vectors are NOT reference
output ("parity unpinned", see oracle/nbody_oracle.h).  They come from a second,
separately written restatement of the reference in pure Python (IEEE-754 double,
no FMA, strict left-to-right — the same FP model as Go gc/amd64), following
  cmd/body/body.go:114-139,148-225,248-264
  cmd/body/collisioncalc.go:26-186
  cmd/body/body_collection.go:212-233 (reverse-arrival resolve order)
so that the C oracle is checked against something other than itself.  The
transcendentals come from Python's math module (glibc libm), as in the C oracle's default backend
(module attribute `M`; see tests/test_gomath.py for the Go-library variant).

Run:  python tests/golden/make_golden.py   (rewrites golden.json deterministically)
"""
import json
import math
import os
import struct

G = 6.673e-11
NONE, SUBSUME, ELASTIC, FRAGMENT = 0, 1, 2, 3

# Transcendentals of calcElasticCollision.  `math` (glibc) for the committed golden.json, like the C
# oracle's default backend; tests/test_gomath.py swaps in tests/golden/gomath_py.py (the Go standard
# library's algorithms restated) and compares the scenes with the C oracle's Go backend bit for bit.
M = math


def hx(v):
    return struct.pack(">d", float(v)).hex()


class Body:
    def __init__(self, x, y, z, vx, vy, vz, mass, radius, behavior=ELASTIC, exists=True):
        self.x, self.y, self.z = float(x), float(y), float(z)
        self.vx, self.vy, self.vz = float(vx), float(vy), float(vz)
        self.mass, self.radius = float(mass), float(radius)
        self.behavior = behavior
        self.exists = exists
        self.fragmenting = False
        self.collided = False
        self.r = 1.0
        self.fx = self.fy = self.fz = 0.0


def dist_of(b, o):
    dx = o.x - b.x
    dy = o.y - b.y
    dz = o.z - b.z
    return dx, dy, dz, math.sqrt(dx * dx + dy * dy + dz * dz)


def compute(bodies, i, events):
    b = bodies[i]
    if not b.exists or b.fragmenting:
        return
    b.fx = b.fy = b.fz = 0.0
    for o in bodies:
        if not o.exists:
            continue
        if o is not b and not o.fragmenting:
            dx, dy, dz, dist = dist_of(b, o)
            if b.collided or dist > b.radius + o.radius:
                force = G * b.mass * o.mass / (dist * dist)
                b.fx += force * dx / dist
                b.fy += force * dy / dist
                b.fz += force * dz / dist
    for j, o in enumerate(bodies):
        if b.collided:
            continue
        _, _, _, dist = dist_of(b, o)
        if dist > b.radius + o.radius:
            continue
        if not (dist <= b.radius + o.radius):
            continue
        # canonical event stream: no self pair; a collision event needs a live j
        # (ResolveCollision's Exists gate makes the dead-j ones no-ops); a subsume event does not
        # (body.go:172-186 has no Exists filter and ResolveSubsume, :228-244, has no gate)
        if j == i:
            continue
        ef = (ELASTIC, FRAGMENT)
        if b.behavior in ef and o.behavior in ef:
            if not o.exists:
                continue
            events.append(("collision", i, j, dist))
        elif b.behavior == SUBSUME or o.behavior == SUBSUME:
            if b.radius > o.radius and dist <= b.radius:
                events.append(("subsume", i, j, dist))
            elif o.radius > b.radius and dist <= o.radius:
                events.append(("subsume", j, i, dist))


def calc_elastic(b, o):
    m1, m2, r1, r2 = b.mass, o.mass, b.radius, o.radius
    x1, y1, z1, x2, y2, z2 = b.x, b.y, b.z, o.x, o.y, o.z
    vx1, vy1, vz1, vx2, vy2, vz2 = b.vx, b.vy, b.vz, o.vx, o.vy, o.vz
    r12 = r1 + r2
    m21 = m2 / m1
    x21, y21, z21 = x2 - x1, y2 - y1, z2 - z1
    vx21, vy21, vz21 = vx2 - vx1, vy2 - vy1, vz2 - vz1
    vx_cm = (m1 * vx1 + m2 * vx2) / (m1 + m2)
    vy_cm = (m1 * vy1 + m2 * vy2) / (m1 + m2)
    vz_cm = (m1 * vz1 + m2 * vz2) / (m1 + m2)
    d = math.sqrt(x21 * x21 + y21 * y21 + z21 * z21)
    v = math.sqrt(vx21 * vx21 + vy21 * vy21 + vz21 * vz21)
    if v == 0:
        return None
    x2, y2, z2 = x21, y21, z21
    vx1, vy1, vz1 = -vx21, -vy21, -vz21
    try:
        q = z2 / d
    except ZeroDivisionError:
        q = math.nan  # Go: 0/0 = NaN, no panic for floats
    theta2 = M.acos(q) if not math.isnan(q) else math.nan
    phi2 = 0.0 if (x2 == 0 and y2 == 0) else M.atan2(y2, x2)
    st, ct = (M.sin(theta2), M.cos(theta2)) if not math.isnan(theta2) else (math.nan, math.nan)
    sp, cp = M.sin(phi2), M.cos(phi2)
    vx1r = ct * cp * vx1 + ct * sp * vy1 - st * vz1
    vy1r = cp * vy1 - sp * vx1
    vz1r = st * cp * vx1 + st * sp * vy1 + ct * vz1
    fvz1r = vz1r / v
    if fvz1r > 1:
        fvz1r = 1.0
    elif fvz1r < -1:
        fvz1r = -1.0
    thetav = M.acos(fvz1r) if not math.isnan(fvz1r) else math.nan
    phiv = 0.0 if (vx1r == 0 and vy1r == 0) else M.atan2(vy1r, vx1r)
    dr = d * (M.sin(thetav) if not math.isnan(thetav) else math.nan) / r12
    if thetav > math.pi / 2 or abs(dr) > 1:
        return None
    alpha = M.asin(-dr) if not math.isnan(dr) else math.nan
    beta = phiv
    sbeta, cbeta = (M.sin(beta), M.cos(beta)) if not math.isnan(beta) else (math.nan, math.nan)
    a = M.tan(thetav + alpha) if not math.isnan(thetav + alpha) else math.nan
    dvz2 = 2 * (vz1r + a * (cbeta * vx1r + sbeta * vy1r)) / ((1 + a * a) * (1 + m21))
    vz2r = dvz2
    vx2r = a * cbeta * dvz2
    vy2r = a * sbeta * dvz2
    vz1r = vz1r - m21 * vz2r
    vx1r = vx1r - m21 * vx2r
    vy1r = vy1r - m21 * vy2r
    return dict(
        vx1=ct * cp * vx1r - sp * vy1r + st * cp * vz1r + vx2,
        vy1=ct * sp * vx1r + cp * vy1r + st * sp * vz1r + vy2,
        vz1=ct * vz1r - st * vx1r + vz2,
        vx2=ct * cp * vx2r - sp * vy2r + st * cp * vz2r + vx2,
        vy2=ct * sp * vx2r + cp * vy2r + st * sp * vz2r + vy2,
        vz2=ct * vz2r - st * vx2r + vz2,
        vx_cm=vx_cm, vy_cm=vy_cm, vz_cm=vz_cm)


def resolve(bodies, events):
    for kind, a, b_, _ in reversed(events):
        b, o = bodies[a], bodies[b_]
        if kind == "subsume":
            tm, om = b.mass, o.mass
            b.mass = tm + om
            o.mass = 0.0
            o.exists = False
            continue
        if not b.exists or not o.exists:
            continue
        if b.behavior == ELASTIC and o.behavior in (ELASTIC, FRAGMENT):
            r = calc_elastic(b, o)
            if r is None:
                continue
            # no Fragment bodies in the golden scenes → doElastic
            b.vx = (r["vx1"] - r["vx_cm"]) * b.r + r["vx_cm"]
            b.vy = (r["vy1"] - r["vy_cm"]) * b.r + r["vy_cm"]
            b.vz = (r["vz1"] - r["vz_cm"]) * b.r + r["vz_cm"]
            o.vx = (r["vx2"] - r["vx_cm"]) * b.r + r["vx_cm"]
            o.vy = (r["vy2"] - r["vy_cm"]) * b.r + r["vy_cm"]
            o.vz = (r["vz2"] - r["vz_cm"]) * b.r + r["vz_cm"]
            b.collided = True
            o.collided = True


def update(b, ts, R):
    if not b.exists:
        return
    if not b.collided:
        b.vx += ts * b.fx / b.mass
        b.vy += ts * b.fy / b.mass
        b.vz += ts * b.fz / b.mass
    b.x += ts * b.vx
    b.y += ts * b.vy
    b.z += ts * b.vz
    b.collided = False
    b.r = R
    if math.isnan(b.x) or math.isnan(b.y) or math.isnan(b.z):
        b.exists = False


def run_scene(name, bodies, ts, R, steps=1):
    init = [dict(x=hx(b.x), y=hx(b.y), z=hx(b.z), vx=hx(b.vx), vy=hx(b.vy), vz=hx(b.vz),
                 mass=hx(b.mass), radius=hx(b.radius), behavior=b.behavior, exists=b.exists)
            for b in bodies]
    out_steps = []
    for _ in range(steps):
        events = []
        for i in range(len(bodies)):
            compute(bodies, i, events)
        forces = [[hx(b.fx), hx(b.fy), hx(b.fz)] for b in bodies]
        resolve(bodies, events)
        for b in bodies:
            update(b, ts, R)
        out_steps.append(dict(
            forces=forces,
            events=[[k, a, b_, hx(d)] for k, a, b_, d in events],
            state=[dict(x=hx(b.x), y=hx(b.y), z=hx(b.z), vx=hx(b.vx), vy=hx(b.vy), vz=hx(b.vz),
                        mass=hx(b.mass), exists=b.exists) for b in bodies]))
    return dict(name=name, ts=hx(ts), R=hx(R), init=init, steps=out_steps)


class Lcg:
    """Tiny deterministic generator (no dependence on numpy/random versions)."""

    def __init__(self, seed):
        self.s = seed & ((1 << 64) - 1)

    def u(self):
        self.s = (self.s * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        return ((self.s >> 11) & ((1 << 53) - 1)) / float(1 << 53)


def scenes():
    out = []
    # KAT-1: cmd/runner/workpool_test.go:41-56
    out.append(run_scene("kat1_wpcompute",
                         [Body(1, 1, 1, 0, 0, 0, 1, 1), Body(22, 22, 22, 0, 0, 0, 1, 1)], 1.0, 1.0))
    # KAT-2: SimTest scene, cmd/sim/simgen.go:387-404
    out.append(run_scene("kat2_simtest", [
        Body(20000, 20000, 20000, -3, -3, -5, 1, 500, SUBSUME),
        Body(0, 0, 0, 0, 0, 0, 9e29, 60),
        Body(-350, 350, 0, 530000000, -500000000, 0, 9e29, 60),
        Body(350, 350, 0, -530000000, -500000000, 0, 9e29, 60)], 1e-9, 1.0, steps=3))
    # KAT-3 head-on equal mass
    out.append(run_scene("kat3_headon",
                         [Body(0, 0, 0, 1, 0, 0, 1, 1), Body(1.5, 0, 0, -1, 0, 0, 1, 1)], 1e-3, 1.0))
    # KAT-4 oblique
    out.append(run_scene("kat4_oblique",
                         [Body(0, 0, 0, 3, 2, 1, 2, 1), Body(1.2, 1.1, 0.9, -1, 0.5, -2, 3, 1.5)], 1e-3, 1.0))
    # KAT-5 separating (mirrored event is a no-op)
    out.append(run_scene("kat5_separating",
                         [Body(0, 0, 0, -1, 0, 0, 1, 1), Body(1.5, 0, 0, 1, 0, 0, 1, 1)], 1e-3, 1.0))
    # KAT-6 coincident centres (cmd/body/body_collection_test.go:319-344): NaN cull
    out.append(run_scene("kat6_coincident",
                         [Body(500, 500, 500, 1, 2, 3, 5, 2), Body(500, 500, 500, 3, 2, 1, 7, 2),
                          Body(900, 0, 0, 0, 0, 0, 4, 1)], 1e-3, 1.0, steps=2))
    # restitution 0.5 takes effect from the second step (Update sets r=R after the first)
    out.append(run_scene("restitution_half",
                         [Body(0, 0, 0, 1, 0.1, 0, 2, 1), Body(1.9, 0, 0, -1, 0, 0.2, 1, 1)], 1e-4, 0.5, steps=3))
    # dense mixed cloud: 48 bodies, overlapping, mixed behaviours, one dead, one subsume sun
    g = Lcg(12345)
    bodies = []
    for i in range(48):
        beh = ELASTIC
        if i % 11 == 3:
            beh = NONE
        b = Body((g.u() - 0.5) * 22, (g.u() - 0.5) * 22, (g.u() - 0.5) * 22,
                 (g.u() - 0.5) * 2e3, (g.u() - 0.5) * 2e3, (g.u() - 0.5) * 2e3,
                 1e12 * (0.5 + g.u()), 1.0 + 3.0 * g.u(), beh)
        bodies.append(b)
    bodies[7].exists = False
    bodies[7].mass = 0.0
    bodies[20] = Body(0, 0, 0, -3, -3, -5, 5e13, 9, SUBSUME)
    out.append(run_scene("dense_mixed_48", bodies, 1e-3, 0.9, steps=4))
    # bodies removed at the cycle top (mod-body exists=false keeps the mass; RemoveBodies zeroes it) stay
    # in the array until Cycle: the collision sweep still visits them and ResolveSubsume is not gated
    # (body.go:172-186,228-244).  0: deleted Subsume body with the larger radius -> swallows live 1;
    # 2: live subsumer, 3: deleted small elastic body inside it that kept its mass -> 2 gains it;
    # 4: deleted elastic body overlapping live elastic 5 -> collision event is a no-op, not queued
    dj = [Body(0, 0, 0, 0, 0, 0, 7e10, 5.0, SUBSUME), Body(1, 1, 0, 2, 0, 0, 1e10, 1.0),
          Body(100, 0, 0, 0, 1, 0, 4e10, 6.0, SUBSUME), Body(101, 2, 0, 0, 0, 0, 3e10, 1.0),
          Body(200, 0, 0, 0, 0, 1, 2e10, 2.0), Body(201, 0, 0, -1, 0, 0, 2e10, 2.0)]
    dj[0].exists = False
    dj[3].exists = False
    dj[4].exists = False
    dj[4].mass = 0.0
    out.append(run_scene("dead_j_subsume", dj, 1e-3, 1.0, steps=2))
    # touching pair: dist == r1+r2 exactly (predicate edge: collision, no force)
    out.append(run_scene("touching_exact",
                         [Body(0, 0, 0, 0.5, 0, 0, 1e10, 1.5), Body(3, 0, 0, -0.5, 0, 0, 1e10, 1.5),
                          Body(0, 3.0000000000000004, 0, 0, 0, 0, 1e10, 1.5)], 1e-3, 1.0))
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as f:
        json.dump(scenes(), f, indent=0, separators=(",", ":"))
    print("wrote", path)
