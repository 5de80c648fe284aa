"""Checkpoint / resume through the CSV channel (the reference can seed state from CSV but
never writes it back; §8f row 1 asks for a dump that doubles as checkpoint): 10 + 10 cycles through a
final-state dump must equal 20 cycles in one go, bit for bit (%.17g round-trips fp64).  The CSV alone is enough
for R = 1 without fragmentation; the state sidecar (--dump-final-state / --resume-state) carries the rest."""
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


def test_resume_with_fragmentation_in_flight_and_R(tmp_path):
    """What the CSV cannot carry — ids, names, fragmenting + fragInfo, the forces a fragmenting body keeps, per-body
    r, the id generator, the cycle counter (it seeds where fragments appear) and the runner's R — travels in the
    sidecar: 3 + 3 cycles ≡ 6 cycles bit for bit while a body is shedding ~100 fragments per cycle."""
    exe = _build.build_host()["nbody_server"]
    f = {k: str(tmp_path / k) for k in ("start.csv", "mid.csv", "mid.state", "a.csv", "a.state", "b.csv", "b.state")}
    rows = ["0,0,0,0,0,0,9e20,100,false,subsume,red,0,0", "50,0,0,0,0,0,1e10,5,false,elastic,blue,0,0",
            "1000,0,0,1e9,0,0,1e12,10,false,elastic,green,0,0",
            "1015,0,0,-1e9,2e8,0,1e12,10,false,fragment,yellow,0.01,100"]
    cloud = clouds.uniform_cube(300, 400.0, 2.0, 1e12, vmax=1e9, seed=12)
    clouds.write_csv(f["start.csv"], cloud)
    with open(f["start.csv"], "a") as fh:
        fh.write("\n".join(rows) + "\n")

    def run(csv, iters, out_csv, out_state, extra):
        r = subprocess.run([exe, "--csv", csv, "--bodies=100000", f"--iterations={iters}", "--collision=elastic",
                            "--no-render", "--no-barnes-hut", "--scaling", "1e-12", f"--dump-final-csv={out_csv}",
                            f"--dump-final-state={out_state}", *extra], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]

    run(f["start.csv"], 6, f["a.csv"], f["a.state"], ["--restitution=0.8"])
    run(f["start.csv"], 3, f["mid.csv"], f["mid.state"], ["--restitution=0.8"])
    mid = open(f["mid.state"]).read().splitlines()
    assert any(line.split("\t")[6] == "1" for line in mid[1:]), "no body was fragmenting at the checkpoint"
    next_id = lambda path: int(open(path).readline().split("\t")[2])
    assert next_id(f["mid.state"]) > 305                            # fragments already arrived
    run(f["mid.csv"], 3, f["b.csv"], f["b.state"], [f"--resume-state={f['mid.state']}"])
    for kind in ("csv", "state"):
        a, b = (open(f[f"{k}.{kind}"]).read().splitlines() for k in "ab")
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y]
        assert len(a) == len(b) and not diff, (kind, len(a), len(b), len(diff), diff[:3])
    assert next_id(f["b.state"]) > next_id(f["mid.state"]) + 50     # and kept arriving after the resume

    # a sidecar that does not belong to the CSV is refused
    r = subprocess.run([exe, "--csv", f["start.csv"], "--bodies=100000", "--iterations=1", "--no-render",
                        "--no-barnes-hut", f"--resume-state={f['mid.state']}"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "does not match" in r.stderr
