"""In-tree nvcc build of libnbody_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libnbody_b200.so")
SOURCES = ("nb_force.cu", "nb_resolve.cu", "nb_integrate.cu", "nb_api.cu")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # no implicit FMA contraction anywhere: the kernels that restate reference arithmetic must
    # round like Go's unfused expressions; the force fast path uses explicit __fma_rn
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(PKG, "..", "include", "nbody_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("NB_NVCC_EXTRA", "").split()
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", SO, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
