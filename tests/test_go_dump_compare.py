"""The reference pin is one command away: integration/cmd/sim/parity_dump_test.go dumps the real Go path,
tests/parity/compare_go_dump.py compares the dump with the oracle (and the GPU).  No Go toolchain exists
here, so this checks the comparer and the dump format on a dump the oracle writes itself, and that the
comparer does notice a single flipped bit."""
import os
import subprocess
import sys

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
