"""Structure-of-arrays host image of the reference's ``[]*Body``.

Mirrors the numeric part of ``Body`` (cmd/body/body.go:34-52): the eight fp64
state fields the device owns (x, y, z, vx, vy, vz, mass, radius), the private
restitution ``r`` (body.go:46), the fragmentation knobs, the collision
behaviour enum (cmd/globals/globals.go:11-16) and the boolean fields packed in
one flag byte.  Identity (Id/Name/Class/colour) stays with the host
application, exactly as SURVEY.md §8b assigns ownership.
"""
from __future__ import annotations

import numpy as np

# cmd/globals/globals.go:11-16
NONE, SUBSUME, ELASTIC, FRAGMENT = 0, 1, 2, 3
BEHAVIOR_NAMES = ("none", "subsume", "elastic", "fragment")

# flag bits — values are part of the C ABI (include/nbody_b200.h)
F_EXISTS = 0x01
F_FRAGMENTING = 0x02
F_PINNED = 0x04
F_SUN = 0x08
F_TELEMETRY = 0x10
F_COLLIDED = 0x20

F64_FIELDS = ("x", "y", "z", "vx", "vy", "vz", "mass", "radius", "rest", "frag_factor", "frag_step")
U8_FIELDS = ("behavior", "flags")


def parse_collision_behavior(s: str) -> int:
    """globals.ParseCollisionBehavior (cmd/globals/globals.go:39-46): unknown → Elastic."""
    s = s.lower()
    return BEHAVIOR_NAMES.index(s) if s in BEHAVIOR_NAMES else ELASTIC


class BodyArrays:
    """SoA body state on the host. All arrays have length ``n``."""

    def __init__(self, n: int = 0):
        self.n = int(n)
        for f in F64_FIELDS:
            setattr(self, f, np.zeros(self.n, dtype=np.float64))
        self.rest[:] = 1.0  # NewBody: r = 1 (cmd/body/body.go:79)
        self.behavior = np.full(self.n, ELASTIC, dtype=np.uint8)
        self.flags = np.full(self.n, F_EXISTS, dtype=np.uint8)
        self.id = np.arange(self.n, dtype=np.int64)

    # -- construction helpers -------------------------------------------
    @classmethod
    def from_fields(cls, x, y, z, vx, vy, vz, mass, radius, behavior=ELASTIC, flags=F_EXISTS):
        n = len(x)
        b = cls(n)
        for name, val in zip(("x", "y", "z", "vx", "vy", "vz", "mass", "radius"),
                             (x, y, z, vx, vy, vz, mass, radius)):
            getattr(b, name)[:] = np.asarray(val, dtype=np.float64)
        b.behavior[:] = behavior
        b.flags[:] = flags
        return b

    def copy(self) -> "BodyArrays":
        c = BodyArrays(0)
        c.n = self.n
        for f in F64_FIELDS + U8_FIELDS + ("id",):
            setattr(c, f, getattr(self, f).copy())
        return c

    def append(self, other: "BodyArrays") -> None:
        for f in F64_FIELDS + U8_FIELDS + ("id",):
            setattr(self, f, np.concatenate([getattr(self, f), getattr(other, f)]))
        self.n += other.n

    def take(self, idx) -> "BodyArrays":
        c = BodyArrays(0)
        idx = np.asarray(idx)
        for f in F64_FIELDS + U8_FIELDS + ("id",):
            setattr(c, f, np.ascontiguousarray(getattr(self, f)[idx]))
        c.n = len(c.x)
        return c

    @property
    def exists(self):
        return (self.flags & F_EXISTS) != 0
