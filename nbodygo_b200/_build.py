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


class _BuildLock:
    """One builder at a time per checkout: several ranks of one job (torchrun) may all find the library stale
    (a copy that did not keep the mtimes) — they must not write the same file concurrently."""

    def __enter__(self):
        import fcntl
        self.f = open(os.path.join(PKG, ".build.lock"), "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return SO
    with _BuildLock():
        if not force and not _stale():      # another rank built it while this one waited
            return SO
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        extra = os.environ.get("NB_NVCC_EXTRA", "").split()
        tmp = SO + f".tmp{os.getpid()}"
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, SO)                     # atomic: a reader never sees a half-written library
        if verbose:
            print(r.stderr)
    return SO


def build_variant_full(name: str, flags: str) -> str:
    """Development only: the whole library compiled with extra -D knobs (for knobs that touch nb_internal.cuh)."""
    vdir = os.path.join(PKG, "variants")
    os.makedirs(vdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    so = os.path.join(vdir, f"libnbody_b200_{name}.so")
    cmd = [nvcc, *NVCC_FLAGS, *flags.split(), "-o", so, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return so


def build_variant(name: str, force_flags: str) -> str:
    """Development only (tools/k1_hw_variants.py): a copy of the library whose nb_force.cu is compiled with extra
    -D knobs, as nbodygo_b200/variants/libnbody_b200_<name>.so.  The other translation units are compiled once."""
    vdir = os.path.join(PKG, "variants")
    os.makedirs(vdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    base = [nvcc, *[f for f in NVCC_FLAGS if f != "-shared"], "-c"]
    objs = []
    for src in SOURCES:
        if src == "nb_force.cu":
            continue
        obj = os.path.join(vdir, src.replace(".cu", ".o"))
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(
                os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if not os.path.isdir(os.path.join(CSRC, f))):
            r = subprocess.run([*base, "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        objs.append(obj)
    fobj = os.path.join(vdir, f"nb_force_{name}.o")
    r = subprocess.run([*base, *force_flags.split(), "-o", fobj, os.path.join(CSRC, "nb_force.cu")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    so = os.path.join(vdir, f"libnbody_b200_{name}.so")
    r = subprocess.run([nvcc, "-shared", "-o", so, fobj, *objs, "-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return so


HOST_DIR = os.path.join(CSRC, "host")
BIN_DIR = os.path.join(PKG, "bin")
HOST_LIB_SOURCES = ("host_core.cc", "host_runner.cc", "host_sim.cc")
HOST_BINARIES = {"host_tests": "host_tests.cc", "nbody_server": "nbody_server.cc"}


def build_host(force: bool = False) -> dict:
    """g++ build of the C++ host mirror (cmd/body, cmd/runner, cmd/sim restated) and its two
    executables, linked against libnbody_b200.so (rpath $ORIGIN/..)."""
    build()
    os.makedirs(BIN_DIR, exist_ok=True)
    out = {}
    deps = [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR)] + [SO]
    newest = max(os.path.getmtime(d) for d in deps)
    cxx = os.environ.get("CXX", "g++")
    with _BuildLock():
        for name, main_src in HOST_BINARIES.items():
            exe = os.path.join(BIN_DIR, name)
            out[name] = exe
            if not force and os.path.exists(exe) and os.path.getmtime(exe) >= newest:
                continue
            tmp = exe + f".tmp{os.getpid()}"
            cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", tmp,
                   *[os.path.join(HOST_DIR, s) for s in HOST_LIB_SOURCES], os.path.join(HOST_DIR, main_src),
                   "-L" + PKG, "-lnbody_b200", "-Wl,-rpath,$ORIGIN/..", "-lpthread"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
            os.replace(tmp, exe)
    return out


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
