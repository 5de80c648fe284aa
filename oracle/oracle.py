"""ctypes wrapper around oracle/libnbody_oracle.so.

TEST INFRASTRUCTURE ONLY — see oracle/nbody_oracle.h.  PARITY UNPINNED.
Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs;
never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OPT_SELF_PAIRS = 0x1
OPT_DEAD_J = 0x2
OPT_SINGLE_SWEEP = 0x4
OPT_DEAD_J_SUBSUME = 0x8
# the canonical event stream: i != j; collision events need a live j (dead-j ones are
# no-ops in ResolveCollision); subsume events do not (ResolveSubsume has no Exists gate)
OPT_CANONICAL = OPT_DEAD_J_SUBSUME

EV_COLLISION, EV_SUBSUME, EV_FRAGMENT, EV_FRAG_INIT = 0, 1, 2, 3

_DP = C.POINTER(C.c_double)
_U8P = C.POINTER(C.c_uint8)


class _OrcBodies(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(f, _DP) for f in (
        "x", "y", "z", "vx", "vy", "vz", "mass", "radius", "rest", "frag_factor", "frag_step",
        "fx", "fy", "fz")] + [("behavior", _U8P), ("flags", _U8P)]


class OrcEvent(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32),
                ("dist", C.c_double), ("f1", C.c_double), ("f2", C.c_double)]


EVENT_DTYPE = np.dtype([("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("_pad", "<i4"),
                        ("dist", "<f8"), ("f1", "<f8"), ("f2", "<f8")])
assert EVENT_DTYPE.itemsize == C.sizeof(OrcEvent)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libnbody_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("nbody_oracle.c", "nbody_oracle.h", "gomath.c", "gomath.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        P = C.POINTER(_OrcBodies)
        EP = C.POINTER(OrcEvent)
        I64P = C.POINTER(C.c_int64)
        L.orc_compute.argtypes = [P, C.c_int64, C.c_int64, C.c_uint32, EP, C.c_int64, I64P]
        L.orc_compute_exact.argtypes = [P, C.c_int64, C.c_int64, _DP, _DP, _DP, _DP]
        L.orc_compute_pool.argtypes = [P, C.c_int, C.c_uint32, EP, C.c_int64, I64P]
        L.orc_compute_slice_timed.argtypes = [P, C.c_int64, C.c_int64, C.c_int, C.c_uint32]
        L.orc_compute_slice_timed.restype = C.c_int64
        L.orc_process_mods.argtypes = [P, EP, C.c_int64, EP, C.c_int64, I64P]
        L.orc_calc_elastic.argtypes = [P, C.c_int64, C.c_int64, _DP]
        L.orc_calc_elastic.restype = None
        L.orc_update.argtypes = [P, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                 C.POINTER(C.c_float), _U8P]
        L.orc_cycle_compact.argtypes = [P, I64P]
        L.orc_cycle_compact.restype = C.c_int64
        L.orc_set_math.argtypes = [C.c_int]
        L.orc_get_math.argtypes = []
        for f in ("go_sin", "go_cos", "go_tan", "go_atan", "go_asin", "go_acos"):
            getattr(L, f).argtypes = [C.c_double]
            getattr(L, f).restype = C.c_double
        L.go_atan2.argtypes = [C.c_double, C.c_double]
        L.go_atan2.restype = C.c_double
        _LIB = L
    return _LIB


MATH_LIBM, MATH_GO = 0, 1


def set_math(which: int) -> int:
    """Selects the transcendental backend of calcElasticCollision (process-wide): MATH_LIBM
    (glibc, default) or MATH_GO (the Go standard library's algorithms, oracle/gomath.c).
    Returns the previous backend."""
    L = lib()
    prev = L.orc_get_math()
    if L.orc_set_math(which) != 0:
        raise ValueError(f"unknown math backend {which}")
    return prev


def _dp(a):
    return a.ctypes.data_as(_DP)


class OracleSim:
    """Holds a BodyArrays (mutated in place) plus force arrays; runs oracle calls."""

    def __init__(self, bodies):
        self.b = bodies
        n = bodies.n
        self.fx = np.zeros(n)
        self.fy = np.zeros(n)
        self.fz = np.zeros(n)
        self.events = np.zeros(0, dtype=EVENT_DTYPE)
        self.host_events = np.zeros(0, dtype=EVENT_DTYPE)
        self.render_xyz = None
        self.render_exists = None

    def _struct(self):
        b = self.b
        for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius", "rest", "frag_factor", "frag_step"):
            a = getattr(b, f)
            assert a.dtype == np.float64 and a.flags.c_contiguous and len(a) >= b.n, f
        s = _OrcBodies()
        s.n = b.n
        for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius", "rest", "frag_factor", "frag_step"):
            setattr(s, f, _dp(getattr(b, f)))
        s.fx, s.fy, s.fz = _dp(self.fx), _dp(self.fy), _dp(self.fz)
        s.behavior = b.behavior.ctypes.data_as(_U8P)
        s.flags = b.flags.ctypes.data_as(_U8P)
        return s

    # ---- Body.Compute over a slice ------------------------------------
    def compute(self, i0=0, i1=None, opts=OPT_CANONICAL, workers=None, ev_cap=None):
        b = self.b
        i1 = b.n if i1 is None else i1
        cap = ev_cap if ev_cap is not None else max(4096, 64 * b.n)
        while True:
            ev = np.zeros(cap, dtype=EVENT_DTYPE)
            n_ev = C.c_int64(0)
            s = self._struct()
            evp = ev.ctypes.data_as(C.POINTER(OrcEvent))
            if workers is None:
                rc = lib().orc_compute(C.byref(s), i0, i1, opts, evp, cap, C.byref(n_ev))
            else:
                rc = lib().orc_compute_pool(C.byref(s), workers, opts, evp, cap, C.byref(n_ev))
            if rc == -1 and ev_cap is None:
                cap = int(n_ev.value) + 16
                continue
            if rc not in (0,):
                raise RuntimeError(f"oracle compute failed rc={rc}")
            break
        self.events = ev[: n_ev.value].copy()
        return self.events

    def compute_exact(self, i0=0, i1=None):
        b = self.b
        i1 = b.n if i1 is None else i1
        fx, fy, fz, fn = (np.zeros(b.n) for _ in range(4))
        s = self._struct()
        lib().orc_compute_exact(C.byref(s), i0, i1, _dp(fx), _dp(fy), _dp(fz), _dp(fn))
        return fx, fy, fz, fn

    def time_slice(self, i0, i1, workers, opts=0):
        s = self._struct()
        return lib().orc_compute_slice_timed(C.byref(s), i0, i1, workers, opts)

    # ---- ProcessMods ----------------------------------------------------
    def process_mods(self, events=None):
        ev = self.events if events is None else events
        ev = np.ascontiguousarray(ev)
        out = np.zeros(max(16, 3 * len(ev)), dtype=EVENT_DTYPE)
        n_out = C.c_int64(0)
        s = self._struct()
        rc = lib().orc_process_mods(C.byref(s), ev.ctypes.data_as(C.POINTER(OrcEvent)), len(ev),
                                    out.ctypes.data_as(C.POINTER(OrcEvent)), len(out), C.byref(n_out))
        if rc:
            raise RuntimeError(f"oracle process_mods failed rc={rc}")
        self.host_events = out[: n_out.value].copy()
        return self.host_events

    def calc_elastic(self, a, b):
        out = np.zeros(10)
        s = self._struct()
        lib().orc_calc_elastic(C.byref(s), a, b, _dp(out))
        return bool(out[0]), out[1:4].copy(), out[4:7].copy(), out[7:10].copy()

    # ---- Update ---------------------------------------------------------
    def update(self, time_scaling, R, i0=0, i1=None):
        b = self.b
        i1 = b.n if i1 is None else i1
        self.render_xyz = np.zeros((b.n, 3), dtype=np.float32)
        self.render_exists = np.zeros(b.n, dtype=np.uint8)
        s = self._struct()
        lib().orc_update(C.byref(s), i0, i1, time_scaling, R,
                         self.render_xyz.ctypes.data_as(C.POINTER(C.c_float)),
                         self.render_exists.ctypes.data_as(_U8P))

    def cycle_compact(self):
        b = self.b
        m = np.zeros(b.n, dtype=np.int64)
        s = self._struct()
        n_new = lib().orc_cycle_compact(C.byref(s), m.ctypes.data_as(C.POINTER(C.c_int64)))
        keep = m[:n_new]
        b.id = b.id[keep].copy()
        for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius", "rest", "frag_factor",
                  "frag_step", "behavior", "flags"):
            setattr(b, f, getattr(b, f)[:n_new].copy())
        self.fx, self.fy, self.fz = self.fx[:n_new].copy(), self.fy[:n_new].copy(), self.fz[:n_new].copy()
        b.n = int(n_new)
        return keep

    def step(self, time_scaling, R, opts=OPT_CANONICAL, workers=None):
        """compute → ProcessMods → Update (cmd/runner/computation-runner.go:297-320)."""
        self.compute(opts=opts, workers=workers)
        self.process_mods()
        self.update(time_scaling, R)

    # canonical pair set the GPU path emits: collision events, i != j, both exist (OPT_CANONICAL)
    def collision_pairs(self):
        ev = self.events
        m = ev["kind"] == EV_COLLISION
        return np.stack([ev["a"][m], ev["b"][m]], axis=1).astype(np.int32)
