"""Build-time guard on K1's hot loops (CPU: reads the SASS of the built library with cuobjdump).

The loops are written against a measured issue model (profiles/r1_summary.md); a source change
that makes ptxas spill, re-introduce per-pair MOVs or lose the tile pipeline would cost several
per cent of the headline without failing any numerical test — and without a GPU nobody would see
it.  This pins the instruction budget of the production instantiations instead.
"""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not installed")

PROD = "k_forceILi4ELi128ELi1ELi1ELi256"     # R = 4, 128 threads, 256-body tiles (16,384 <= n < 786,432)
PROD_HUGE = "k_forceILi4ELi128ELi1ELi1ELi512"  # the same with 512-body tiles (n >= 786,432: C4 runs these)


@pytest.fixture(scope="module", params=[(PROD, "ELi256ELi"), (PROD_HUGE, "ELi512ELi")], ids=["tiles256", "tiles512"])
def loops(request):
    import sass_model
    from nbodygo_b200 import _build
    prod, tag = request.param
    rows = sass_model.hot_loops(_build.build(), prod)
    by = {}
    for name, addr, n, hist, n_fp64, three, cycles in rows:
        mode = int(name.split(tag)[1][0])          # FORCE_ALL 0, FORCE_MIXED 1, FORCE_UNI 2
        by[(mode, bool(hist.get("SEL", 0)))] = dict(n=n, hist=hist, fp64=n_fp64, three=three, cycles=cycles)
    return by


def test_every_production_pass_has_its_two_fast_loops(loops):
    assert set(loops) == {(m, s) for m in (0, 1, 2) for s in (False, True)}


@pytest.mark.parametrize("mode,fp64,lds,max_instr,max_cycles", [
    (0, 128, 4, 150, 290),    # per-body masses: 16 FP64 instructions per pair
    (1, 128, 4, 150, 290),
    (2, 120, 3, 141, 269),    # uniform-mass pass: 15 per pair, no LDS of the masses
])
def test_fast_loop_budget(loops, mode, fp64, lds, max_instr, max_cycles):
    """Round 2: the constant 15/8 stays in a register pair (one IMAD.MOV less per trip; measured -0.55 % on the
    uniform-mass launch).  What the hardware measurements of round 2 showed (profiles/r2_k1_variants.txt): the
    instruction COUNT is what matters — every non-FP64 instruction costs about 1.5 issue cycles with two warps per
    scheduler — while the three-register reads the round-1 model charged a cycle for are free here.  So the budget
    pins counts; `cycles` (the round-1 model) is kept as a loose cap only."""
    L = loops[(mode, False)]
    h = L["hist"]
    assert L["fp64"] == fp64                      # 8 pairs per iteration
    assert h.get("MUFU") == 8 and h.get("LDS") == lds and h.get("VIMNMX3") == 4
    assert h.get("IMAD", 0) <= 1                  # no per-pair MOVs, no rematerialised constant
    assert not any(k in h for k in ("LDL", "STL", "LDG", "CALL", "BAR", "SEL", "ISETP"))
    assert L["n"] <= max_instr and L["cycles"] <= max_cycles, (L["n"], L["cycles"])


def test_self_tile_loops_are_branch_free_too(loops):
    for mode, fp64 in ((0, 128), (1, 128), (2, 120)):
        L = loops[(mode, True)]
        assert L["fp64"] == fp64 and L["hist"].get("BRA") == 1
        assert not any(k in L["hist"] for k in ("LDL", "STL", "CALL", "BAR"))


def test_production_kernels_do_not_spill():
    import subprocess
    from nbodygo_b200 import _build
    out = subprocess.run(["cuobjdump", "-res-usage", _build.build()], capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    seen = 0
    for i, line in enumerate(lines):
        if PROD in line or PROD_HUGE in line:
            use = lines[i + 1]
            assert "STACK:0" in use and "LOCAL:0" in use, use
            seen += 1
    assert seen == 6
