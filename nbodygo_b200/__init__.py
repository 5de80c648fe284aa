"""nbodygo_b200 — B200-native replacement for the nbodygo work pool (one hot path).

Only what the per-cycle compute path needs lives here: the CUDA kernels and C ABI
(``csrc/``, ``include/nbody_b200.h``), the ctypes binding (``capi``), the SoA host
image of ``[]*Body`` (``bodies``) and the seeded input clouds / CSV channel (``clouds``).
"""
from .bodies import BodyArrays  # noqa: F401

__all__ = ["BodyArrays"]
