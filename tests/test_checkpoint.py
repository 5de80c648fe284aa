"""Checkpoint / resume through the CSV channel (the reference can seed state from CSV but
never writes it back; §8f row 1 asks for a dump that doubles as checkpoint): 10 + 10 cycles through a
final-state dump must equal 20 cycles in one go, bit for bit (%.17g round-trips fp64)."""
import subprocess

import pytest

from nbodygo_b200 import _build, clouds

pytestmark = pytest.mark.gpu


def _run(exe, args):
    r = subprocess.run([exe, *args, "--collision=elastic", "--no-render", "--no-barnes-hut", "--scaling", "1e-9"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_resume_from_final_state_dump(tmp_path):
    exe = _build.build_host()["nbody_server"]
    start, mid, end_a, end_b = (str(tmp_path / f) for f in ("start.csv", "mid.csv", "end_a.csv", "end_b.csv"))
    clouds.write_csv(start, clouds.config("C1", n=801))
    _run(exe, ["--csv", start, "--bodies=801", "--iterations=20", f"--dump-final-csv={end_a}"])
    _run(exe, ["--csv", start, "--bodies=801", "--iterations=10", f"--dump-final-csv={mid}"])
    _run(exe, ["--csv", mid, "--bodies=801", "--iterations=10", f"--dump-final-csv={end_b}"])
    a, b = open(end_a).read(), open(end_b).read()
    assert len(a.splitlines()) == 802 and a == b
