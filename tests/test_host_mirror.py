"""Runs the C++ host mirror's test executable (nbodygo_b200/csrc/host/host_tests.cc), which
restates the reference's own Go unit tests (cmd/body/*_test.go, cmd/runner/*_test.go) against
the C++ Body / BodyCollection / ResultQueueHolder / ComputationRunner built on the C ABI."""
import subprocess

import pytest

from nbodygo_b200 import _build


def _run(mode, timeout=600):
    exe = _build.build_host()["host_tests"]
    r = subprocess.run([exe, mode], capture_output=True, text=True, timeout=timeout)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    return r


def test_host_unit_tests_cpu():
    r = _run("cpu")
    assert r.returncode == 0, r.stdout[-3000:]
    assert "19 tests, 0 failures" in r.stdout


def test_runner_refuses_to_start_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = _run("nodevice")
    assert r.returncode == 0, r.stdout[-3000:]


@pytest.mark.gpu
def test_host_unit_tests_gpu():
    r = _run("gpu")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "7 tests, 0 failures" in r.stdout


@pytest.mark.gpu
def test_headless_server_runs_csv_identical_input(tmp_path):
    """nbody_server (headless twin of cmd/server) fed the CSV the Go server would be fed."""
    from nbodygo_b200 import clouds
    exe = _build.build_host()["nbody_server"]
    csv = tmp_path / "c.csv"
    clouds.write_csv(str(csv), clouds.config("C1", n=501))
    r = subprocess.run([exe, "--csv", str(csv), "--bodies=501", "--collision=elastic", "--no-render",
                        "--no-barnes-hut", "--iterations=20", "--scaling", "1e-9"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "computations: 20" in r.stdout and "frames per second" in r.stdout
