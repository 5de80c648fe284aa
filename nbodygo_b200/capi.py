"""ctypes binding of libnbody_b200.so (include/nbody_b200.h).

This is what the Go side would reach through cgo; Python only
plays the host application here.  There is no fallback: if the library is not
built, or no CUDA device is present, calls raise ``NbError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build
from .bodies import BodyArrays

NB_OK = 0
NB_ERR_INVALID, NB_ERR_CUDA, NB_ERR_CAPACITY, NB_ERR_PAIR_OVERFLOW, NB_ERR_COMM, NB_ERR_NO_DEVICE = -1, -2, -3, -4, -5, -6

STEP_COLLISIONS = 0x01
STEP_NO_RESOLVE = 0x02
STEP_NO_INTEGRATE = 0x04
STEP_ASYNC = 0x08
STEP_PHASE_TIMINGS = 0x10   # ms_prep .. ms_integrate (costs a few us of a small cycle)
STEP_DEFAULT = STEP_COLLISIONS

EV_COLLISION, EV_SUBSUME, EV_FRAGMENT, EV_FRAG_INIT = 0, 1, 2, 3

COMM_SINGLE, COMM_PEER_PUSH, COMM_NCCL = 0, 1, 2
COMM_MODE_NAMES = {COMM_SINGLE: "single", COMM_PEER_PUSH: "peer_push", COMM_NCCL: "nccl"}

# every symbol include/nbody_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = (
    "nb_create", "nb_destroy", "nb_last_error", "nb_abi_version", "nb_upload", "nb_patch", "nb_append",
    "nb_compact", "nb_count", "nb_step", "nb_sync", "nb_download_state", "nb_download_render", "nb_render_buffers",
    "nb_get_forces", "nb_get_pairs", "nb_get_host_events", "nb_comm_unique_id", "nb_comm_init",
    "nb_shard_range", "nb_plan", "nb_measure_fp64_peak", "nb_launch_count",
    "nb_graph_stats", "nb_upload_shard", "nb_download_state_range", "nb_download_render_range", "nb_comm_mode",
    "nb_set_forces", "nb_get_cycle_top_positions",
)


class NbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnbody_b200 error {code}: {msg}")
        self.code = code


class StepResult(C.Structure):
    _fields_ = [("n_bodies", C.c_int64), ("n_pairs", C.c_int64), ("n_host_events", C.c_int64),
                ("n_resolved", C.c_int64), ("n_culled", C.c_int64), ("n_dead", C.c_int64),
                ("n_subsumed", C.c_int64),
                ("resolve_rounds", C.c_int32), ("pair_overflow", C.c_int32),
                ("ms_total", C.c_float), ("ms_prep", C.c_float), ("ms_force", C.c_float),
                ("ms_exchange", C.c_float), ("ms_resolve", C.c_float), ("ms_integrate", C.c_float)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


EVENT_DTYPE = np.dtype([("kind", "<i4"), ("a", "<i4"), ("b", "<i4"), ("applied", "<i4"),
                        ("dist", "<f8"), ("f1", "<f8"), ("f2", "<f8")])

_DP = C.POINTER(C.c_double)
_U8P = C.POINTER(C.c_uint8)
_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)
_LIB = None


def load(path: str | None = None):
    """Loads (building if stale) the shared library and declares the signatures."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    # NB_LIBRARY_PATH: development override (tools/k1_hw_variants.py times alternative builds of the library)
    so = path or os.environ.get("NB_LIBRARY_PATH") or _build.build()
    L = C.CDLL(so, mode=C.RTLD_GLOBAL)
    H = C.c_void_p
    state_in = [_DP] * 11 + [_U8P] * 2
    L.nb_create.argtypes = [C.c_int, C.c_int64, C.c_int64, C.POINTER(H)]
    L.nb_destroy.argtypes = [H]
    L.nb_last_error.argtypes = [H]
    L.nb_last_error.restype = C.c_char_p
    L.nb_abi_version.argtypes = []
    L.nb_upload.argtypes = [H, C.c_int64] + state_in
    L.nb_patch.argtypes = [H, C.c_int64, C.c_int64] + state_in
    L.nb_upload_shard.argtypes = [H, C.c_int64, C.c_int64, C.c_int64] + state_in
    L.nb_download_state_range.argtypes = [H, C.c_int64, C.c_int64] + [_DP] * 9 + [_U8P] * 2
    L.nb_download_render_range.argtypes = [H, C.c_int64, C.c_int64, C.POINTER(C.c_float), _U8P]
    L.nb_comm_mode.argtypes = [H, C.POINTER(C.c_int)]
    L.nb_append.argtypes = [H, C.c_int64, C.c_double] + [_DP] * 10 + [_U8P] * 2
    L.nb_compact.argtypes = [H, _I64P, _I64P, C.c_int64]
    L.nb_count.argtypes = [H, _I64P]
    L.nb_step.argtypes = [H, C.c_double, C.c_double, C.c_uint32, C.POINTER(StepResult)]
    L.nb_sync.argtypes = [H, C.POINTER(StepResult)]
    L.nb_download_state.argtypes = [H] + [_DP] * 9 + [_U8P] * 2
    L.nb_download_render.argtypes = [H, C.POINTER(C.c_float), _U8P]
    L.nb_render_buffers.argtypes = [H, C.POINTER(C.POINTER(C.c_float)), C.POINTER(_U8P)]
    L.nb_get_forces.argtypes = [H, _DP, _DP, _DP]
    L.nb_set_forces.argtypes = [H, C.c_int64, C.c_int64, _DP, _DP, _DP]
    L.nb_get_cycle_top_positions.argtypes = [H, C.c_int64, C.c_int64, _DP, _DP, _DP]
    L.nb_get_pairs.argtypes = [H, _I32P, _I32P, C.c_int64, _I64P]
    L.nb_get_host_events.argtypes = [H, C.c_void_p, C.c_int64, _I64P]
    L.nb_comm_unique_id.argtypes = [C.c_void_p]
    L.nb_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    L.nb_shard_range.argtypes = [H, _I64P, _I64P]
    L.nb_plan.argtypes = [C.c_int64, C.c_int, C.c_int, _I64P, _I64P, _I32P, _I32P]
    L.nb_measure_fp64_peak.argtypes = [C.c_int, C.c_int, _DP, C.POINTER(C.c_float)]
    L.nb_launch_count.argtypes = [H, _I64P]
    L.nb_graph_stats.argtypes = [H, _I64P, _I64P]
    if path is None:
        _LIB = L
    return L


def _p(a, typ=_DP):
    if a is None:
        return None
    return a.ctypes.data_as(typ)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


class Sim:
    """One device-resident body collection (one GPU)."""

    def __init__(self, capacity: int, device: int = 0, pair_capacity: int = 0):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.nb_create(device, capacity, pair_capacity, C.byref(self.h))
        if rc:
            raise NbError(rc, (self.L.nb_last_error(None) or b"").decode())
        self.capacity = capacity
        self.device = device

    def _chk(self, rc):
        if rc:
            raise NbError(rc, (self.L.nb_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.L.nb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state sync ------------------------------------------------------
    def upload(self, b: BodyArrays):
        arrs = [_f64(getattr(b, f)) for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius", "rest",
                                              "frag_factor", "frag_step")]
        beh, fl = _u8(b.behavior), _u8(b.flags)
        self._chk(self.L.nb_upload(self.h, b.n, *[_p(a) for a in arrs], _p(beh, _U8P), _p(fl, _U8P)))

    def upload_raw(self, n, x, y, z, vx, vy, vz, mass, radius, rest=None, ff=None, fs=None, behavior=None, flags=None):
        arrs = [_f64(a) for a in (x, y, z, vx, vy, vz, mass, radius, rest, ff, fs)]
        beh, fl = _u8(behavior), _u8(flags)
        self._chk(self.L.nb_upload(self.h, n, *[_p(a) for a in arrs], _p(beh, _U8P), _p(fl, _U8P)))

    def upload_shard(self, n, first, count, x, y, z, vx, vy, vz, mass, radius, rest=None, ff=None, fs=None,
                     behavior=None, flags=None):
        """Collective sharded upload: the arrays hold only this handle's slice [first, first+count) of n bodies."""
        arrs = [_f64(a) for a in (x, y, z, vx, vy, vz, mass, radius, rest, ff, fs)]
        beh, fl = _u8(behavior), _u8(flags)
        for a in arrs + [beh, fl]:
            assert a is None or len(a) == count
        self._chk(self.L.nb_upload_shard(self.h, n, first, count, *[_p(a) for a in arrs], _p(beh, _U8P), _p(fl, _U8P)))

    def patch(self, first, count, **fields):
        names = ("x", "y", "z", "vx", "vy", "vz", "mass", "radius", "rest", "frag_factor", "frag_step")
        arrs = [_f64(fields.get(f)) for f in names]
        beh, fl = _u8(fields.get("behavior")), _u8(fields.get("flags"))
        for a in arrs + [beh, fl]:
            assert a is None or len(a) == count
        self._chk(self.L.nb_patch(self.h, first, count, *[_p(a) for a in arrs], _p(beh, _U8P), _p(fl, _U8P)))

    def append(self, b: BodyArrays, R: float):
        arrs = [_f64(getattr(b, f)) for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius",
                                              "frag_factor", "frag_step")]
        beh, fl = _u8(b.behavior), _u8(b.flags)
        self._chk(self.L.nb_append(self.h, b.n, R, *[_p(a) for a in arrs], _p(beh, _U8P), _p(fl, _U8P)))

    def compact(self):
        n = C.c_int64(0)
        old = np.zeros(max(self.count(), 1), dtype=np.int64)
        self._chk(self.L.nb_compact(self.h, C.byref(n), _p(old, _I64P), len(old)))
        return n.value, old[: n.value].copy()

    def count(self) -> int:
        n = C.c_int64(0)
        self._chk(self.L.nb_count(self.h, C.byref(n)))
        return n.value

    # ---- the cycle -------------------------------------------------------
    def step(self, time_scaling: float, R: float = 1.0, opts: int = STEP_DEFAULT) -> StepResult:
        res = StepResult()
        self._chk(self.L.nb_step(self.h, time_scaling, R, opts, C.byref(res)))
        return res

    def sync(self) -> StepResult:
        res = StepResult()
        self._chk(self.L.nb_sync(self.h, C.byref(res)))
        return res

    # ---- results ---------------------------------------------------------
    def download(self) -> BodyArrays:
        n = self.count()
        b = BodyArrays(n)
        self._chk(self.L.nb_download_state(
            self.h, *[_p(getattr(b, f)) for f in ("x", "y", "z", "vx", "vy", "vz", "mass", "radius", "rest")],
            _p(b.behavior, _U8P), _p(b.flags, _U8P)))
        return b

    def download_into(self, x=None, y=None, z=None, vx=None, vy=None, vz=None):
        self._chk(self.L.nb_download_state(self.h, _p(x), _p(y), _p(z), _p(vx), _p(vy), _p(vz),
                                           None, None, None, None, None))

    def download_range_into(self, first, count, x=None, y=None, z=None, vx=None, vy=None, vz=None):
        self._chk(self.L.nb_download_state_range(self.h, first, count, _p(x), _p(y), _p(z), _p(vx), _p(vy), _p(vz),
                                                 None, None, None, None, None))

    def render_range(self, first, count, xyz, exists):
        self._chk(self.L.nb_download_render_range(self.h, first, count, xyz.ctypes.data_as(C.POINTER(C.c_float)),
                                                  _p(exists, _U8P)))

    def render(self, xyz=None, exists=None):
        n = self.count()
        xyz = np.zeros((n, 3), dtype=np.float32) if xyz is None else xyz
        exists = np.zeros(n, dtype=np.uint8) if exists is None else exists
        self._chk(self.L.nb_download_render(self.h, xyz.ctypes.data_as(C.POINTER(C.c_float)), _p(exists, _U8P)))
        return xyz, exists

    def render_buffers(self):
        """Library-owned pinned (xyz[cap,3] float32, exists[cap] uint8) views filled by every step."""
        px, pe = C.POINTER(C.c_float)(), _U8P()
        self._chk(self.L.nb_render_buffers(self.h, C.byref(px), C.byref(pe)))
        cap = self.capacity
        xyz = np.ctypeslib.as_array(px, shape=(cap, 3))
        ex = np.ctypeslib.as_array(pe, shape=(cap,))
        return xyz, ex

    def forces(self):
        n = self.count()
        fx, fy, fz = np.zeros(n), np.zeros(n), np.zeros(n)
        self._chk(self.L.nb_get_forces(self.h, _p(fx), _p(fy), _p(fz)))
        return fx, fy, fz

    def cycle_top_positions(self, first, count):
        """(x, y, z) the last step started from, for bodies [first, first+count)."""
        x, y, z = np.zeros(count), np.zeros(count), np.zeros(count)
        self._chk(self.L.nb_get_cycle_top_positions(self.h, first, count, _p(x), _p(y), _p(z)))
        return x, y, z

    def set_forces(self, first, count, fx, fy, fz):
        fx, fy, fz = _f64(fx), _f64(fy), _f64(fz)
        self._chk(self.L.nb_set_forces(self.h, first, count, _p(fx), _p(fy), _p(fz)))

    def pairs(self) -> np.ndarray:
        n = C.c_int64(0)
        self._chk(self.L.nb_get_pairs(self.h, None, None, 0, C.byref(n)))
        i = np.zeros(max(n.value, 1), dtype=np.int32)
        j = np.zeros(max(n.value, 1), dtype=np.int32)
        self._chk(self.L.nb_get_pairs(self.h, _p(i, _I32P), _p(j, _I32P), len(i), C.byref(n)))
        return np.stack([i[: n.value], j[: n.value]], axis=1)

    def host_events(self) -> np.ndarray:
        n = C.c_int64(0)
        self._chk(self.L.nb_get_host_events(self.h, None, 0, C.byref(n)))
        ev = np.zeros(max(n.value, 1), dtype=EVENT_DTYPE)
        self._chk(self.L.nb_get_host_events(self.h, ev.ctypes.data, len(ev), C.byref(n)))
        return ev[: n.value]

    # ---- multi-GPU -------------------------------------------------------
    def comm_init(self, rank: int, nranks: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._chk(self.L.nb_comm_init(self.h, rank, nranks, buf))

    def comm_mode(self) -> int:
        """COMM_SINGLE, COMM_PEER_PUSH (kernels store into the peers' replicas) or COMM_NCCL (all-gathers)."""
        m = C.c_int(-1)
        self._chk(self.L.nb_comm_mode(self.h, C.byref(m)))
        return m.value

    def shard_range(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self._chk(self.L.nb_shard_range(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self) -> int:
        n = C.c_int64(0)
        self._chk(self.L.nb_launch_count(self.h, C.byref(n)))
        return n.value

    def graph_stats(self):
        """(graphs captured, cycles replayed from a graph)."""
        c, r = C.c_int64(0), C.c_int64(0)
        self._chk(self.L.nb_graph_stats(self.h, C.byref(c), C.byref(r)))
        return c.value, r.value


def plan(n: int, rank: int = 0, nranks: int = 1):
    """(i0, i1, n_chunks, tiles_per_chunk) — pure host arithmetic, no device needed."""
    L = load()
    i0, i1, nc, tpc = C.c_int64(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
    rc = L.nb_plan(n, rank, nranks, C.byref(i0), C.byref(i1), C.byref(nc), C.byref(tpc))
    if rc:
        raise NbError(rc, "nb_plan: bad arguments")
    return i0.value, i1.value, nc.value, tpc.value


def uniform_chunks(b: BodyArrays):
    """(uniform j-chunks, j-chunks, uniform tiles, tiles) of K1 for this collection — the host-side
    mirror of K0's `tile_muni` rule and K1's per-chunk dispatch (nb_force.cu), for reports and tests:
    a tile is uniform if all its LIVE bodies have one positive finite mass (slots without a live body
    — bodies that do not exist, non-finite positions, the tail of the last tile — are parked far away
    and match any mass; a fragmenting body stays in place with effective mass 0 and breaks it);
    a chunk runs the uniform-mass pass if all its tiles are.  Collections below 16,384 bodies use
    64-body tiles and a single per-body-mass launch (returns 0 uniform chunks for them); from 786,432
    bodies up the tiles hold 512 bodies."""
    from .bodies import F_EXISTS, F_FRAGMENTING
    n = b.n
    _, _, n_chunks, tpc = plan(n)
    tj = 64 if n < 16384 else (512 if n >= 786432 else 256)
    n_tiles = (n + tj - 1) // tj
    with np.errstate(invalid="ignore"):
        finite = (np.abs(b.x) < 1e150) & (np.abs(b.y) < 1e150) & (np.abs(b.z) < 1e150)
    live = np.zeros(n_tiles * tj, dtype=bool)
    live[:n] = ((b.flags & F_EXISTS) != 0) & finite
    m = np.zeros(n_tiles * tj)
    m[:n] = np.where((b.flags & F_FRAGMENTING) == 0, b.mass, 0.0)
    live, m = live.reshape(n_tiles, tj), m.reshape(n_tiles, tj)
    with np.errstate(invalid="ignore"):
        odd = np.any(live & ~((m > 0) & (m < np.inf)), axis=1)
        lo = np.where(live, m, np.inf).min(axis=1)
        hi = np.where(live, m, 0.0).max(axis=1)
    ok = ~odd & ((lo == hi) | (hi == 0.0))
    if tj == 64:
        return 0, n_chunks, int(ok.sum()), n_tiles
    pad = np.ones(n_chunks * tpc, dtype=bool)
    pad[:n_tiles] = ok
    per_chunk = pad.reshape(n_chunks, tpc).all(axis=1)
    return int(per_chunk.sum()), n_chunks, int(ok.sum()), n_tiles


def comm_unique_id() -> bytes:
    L = load()
    buf = C.create_string_buffer(128)
    rc = L.nb_comm_unique_id(buf)
    if rc:
        raise NbError(rc, (L.nb_last_error(None) or b"").decode())
    return buf.raw


def measure_fp64_peak(device: int = 0, iters: int = 4096):
    L = load()
    tf, ms = C.c_double(0), C.c_float(0)
    rc = L.nb_measure_fp64_peak(device, iters, C.byref(tf), C.byref(ms))
    if rc:
        raise NbError(rc, (L.nb_last_error(None) or b"").decode())
    return tf.value, ms.value
