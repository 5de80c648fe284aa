"""The reference pin is one command away: integration/cmd/sim/parity_dump_test.go dumps the real Go path,
tests/parity/compare_go_dump.py compares the dump with the oracle (and the GPU).  No Go toolchain exists
here, so this checks the comparer and the dump format on a dump the oracle writes itself, and that the
comparer does notice a single flipped bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tests", "parity", "compare_go_dump.py")


def test_self_test_pins():
    r = subprocess.run([sys.executable, TOOL, "--self-test", "--quiet"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PINNED" in r.stdout and "NOT PINNED" not in r.stdout
    go_line = [l for l in r.stdout.splitlines() if l.startswith("oracle (Go math backend)")][0]
    assert "events_equal=True" in go_line and "force_bits_equal=True" in go_line and "state_bits_equal=True" in go_line


def test_a_flipped_bit_is_noticed(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "parity"))
    import compare_go_dump as cg
    from nbodygo_b200 import clouds
    from oracle import oracle as orc
    b = clouds.config("C1", n=201)
    csv = str(tmp_path / "in.csv")
    clouds.write_csv(csv, b)
    bodies = clouds.read_csv(csv)
    dump = str(tmp_path / "dump.txt")
    cg.write_dump_from_oracle(dump, bodies, 1e-9, 1.0, 2, orc.MATH_GO)
    assert cg.main(["--csv", csv, "--dump", dump, "--quiet"]) == 0
    lines = open(dump).read().splitlines()
    k = next(i for i, l in enumerate(lines) if l.startswith("F 7 "))
    t = lines[k].split()
    t[2] = "%016x" % (int(t[2], 16) ^ 1)       # one ulp in fx of body 7
    lines[k] = " ".join(t)
    open(dump, "w").write("\n".join(lines) + "\n")
    assert cg.main(["--csv", csv, "--dump", dump, "--quiet"]) == 1
    # dropping an event is noticed as well
    lines = [l for i, l in enumerate(open(dump).read().splitlines())]
    lines[k] = " ".join(t[:2] + ["%016x" % (int(t[2], 16) ^ 1)] + t[3:])
    e = next(i for i, l in enumerate(lines) if l.startswith("E 0 "))
    del lines[e]
    open(dump, "w").write("\n".join(lines) + "\n")
    assert cg.main(["--csv", csv, "--dump", dump, "--quiet"]) == 1


def test_reference_side_patch_applies(tmp_path):
    """integration/patches/nbodygo-gpu.patch (the edits to existing reference files: runner, builder, sim, server
    flags) must apply cleanly to the reference tree.  Only where the tree exists (this container, not the GPU box);
    the Go files next to it cannot be compiled here (no toolchain) — this keeps at least the patch honest."""
    import shutil
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "cmd", "runner")) or shutil.which("patch") is None:
        pytest.skip("reference tree or patch(1) not available")
    patch = os.path.join(ROOT, "integration", "patches", "nbodygo-gpu.patch")
    files = [l[6:].strip() for l in open(patch) if l.startswith("+++ b/")]
    assert "cmd/runner/computation-runner.go" in files and len(files) == 4
    for f in files:
        dst = tmp_path / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(ref, f), dst)
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", patch], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0 and "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout + r.stderr
    # the new files do not collide with anything the reference already has
    for f in ("cmd/runner/gpustepper.go", "cmd/runner/gpustepper_stub.go", "cmd/body/gpu_accessors.go",
              "cmd/sim/parity_dump_test.go"):
        assert os.path.exists(os.path.join(ROOT, "integration", f)) and not os.path.exists(os.path.join(ref, f))
