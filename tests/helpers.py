"""Shared helpers for the parity tests (golden loading, bit views)."""
import json
import os
import struct

import numpy as np

from nbodygo_b200.bodies import BodyArrays, F_EXISTS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")


def unhex(h):
    return struct.unpack(">d", bytes.fromhex(h))[0]


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).view(np.uint64)


def same_bits(a, b):
    """Bit equality, except that any NaN equals any NaN."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def scene_bodies(scene) -> BodyArrays:
    init = scene["init"]
    b = BodyArrays(len(init))
    for i, r in enumerate(init):
        for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius"):
            getattr(b, f)[i] = unhex(r[f])
        b.behavior[i] = r["behavior"]
        b.flags[i] = F_EXISTS if r["exists"] else 0
    return b


def scene_step_arrays(step):
    """Golden post-step state as arrays."""
    st = step["state"]
    out = {f: np.array([unhex(r[f]) for r in st]) for f in ("x", "y", "z", "vx", "vy", "vz", "mass")}
    out["exists"] = np.array([r["exists"] for r in st])
    out["forces"] = np.array([[unhex(h) for h in row] for row in step["forces"]])
    return out
