"""CPU tests of the drop-in boundary: the library builds for sm_100a, loads, and exports
every symbol include/nbody_b200.h declares; without a device it fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

from nbodygo_b200 import _build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    so = _build.build()
    L = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "nbody_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char \*)\s*\*?(nb_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), sym
    assert L.nb_abi_version() == 3


def test_sm100a_sass_and_tma_present():
    import subprocess
    so = _build.build()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass          # 1-D TMA bulk copy of the j-tiles
    assert "MUFU.RSQ64H" in sass     # fp64 rsqrt seed in the force kernel
    assert "DFMA" in sass


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(capi.NbError) as ei:
        capi.Sim(16)
    assert ei.value.code in (capi.NB_ERR_NO_DEVICE, capi.NB_ERR_CUDA)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nbodygo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", "") or f == "clouds.py", f


def test_uniform_chunk_report_mirrors_the_device_rule():
    # host-side mirror of K0's tile rule / K1's chunk dispatch (pure arithmetic + nb_plan, no device)
    import numpy as np
    from nbodygo_b200 import capi, clouds
    from nbodygo_b200.bodies import F_EXISTS
    from nbodygo_b200.bodies import F_FRAGMENTING
    b = clouds.config("C4")
    assert capi.uniform_chunks(b)[:2] == (64, 64)          # the padding of the tail tile matches any mass
    b.flags[5] &= ~np.uint8(F_EXISTS)                      # ... and so does a body that does not exist
    assert capi.uniform_chunks(b)[:2] == (64, 64)
    b.flags[700_000] |= np.uint8(F_FRAGMENTING)            # a fragmenting body stays in place with mass 0
    assert capi.uniform_chunks(b)[:2] == (63, 64)
    c2 = clouds.config("C2")                               # 10 k bodies, masses U[1e24, 1e25]: single launch
    assert capi.uniform_chunks(c2)[0] == 0
    c3 = clouds.config("C3")
    nu, nc, tu, nt = capi.uniform_chunks(c3)
    assert nu == nc and tu == nt                           # 100 k equal masses: every chunk
    c3.mass[::256] *= 1.5
    assert capi.uniform_chunks(c3)[0] == 0 and capi.uniform_chunks(c3)[2] == 0


def test_header_is_plain_c_and_declares_what_the_library_exports(tmp_path):
    """include/nbody_b200.h is the cgo boundary: it must compile as C (not only as C++), and every function it
    declares must be an exported symbol of the built library (and listed in capi.SYMBOLS)."""
    import re
    import shutil
    import subprocess
    from nbodygo_b200 import _build, capi
    if shutil.which("gcc") is None:
        pytest.skip("gcc not installed")
    hdr = os.path.join(ROOT, "include", "nbody_b200.h")
    src = tmp_path / "t.c"
    src.write_text('#include "nbody_b200.h"\nint main(void) { nb_step_result r; nb_event e; (void)r; (void)e; return nb_abi_version() == NB_ABI_VERSION ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only",
                        "-I", os.path.dirname(hdr), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    declared = set(re.findall(r"^(?:int|const char \*)\s*(nb_[a-z0-9_]+)\(", open(hdr).read(), flags=re.M))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    exported = subprocess.run(["nm", "-D", "--defined-only", _build.build()], capture_output=True, text=True).stdout
    for sym in declared:
        assert re.search(rf"\bT {sym}\b", exported), sym
