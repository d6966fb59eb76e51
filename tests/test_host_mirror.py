"""The C++ mirror of the reference's Go interface (requests.Manager / ReplayWorker over the C-ABI): builds on CPU,
runs its KAT + concurrency driver on the GPU."""
import os
import subprocess

import pytest

import agentainer_lab_b200 as A


def test_host_mirror_builds():
    out = A.build_host()
    assert os.path.exists(out)
    src = open(os.path.join(os.path.dirname(out), "requests.hpp")).read()
    for name in ("StoreRequest", "StoreResponse", "GetPendingRequests", "MarkRequestFailed", "class ReplayWorker", "Start(", "Stop("):
        assert name in src      # same method names as internal/requests (requests.go:64,120,197,228; replay_worker.go:36,53)


@pytest.mark.gpu
def test_host_mirror_kats_and_threads():
    out = A.build_host()
    res = subprocess.run([out], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "host mirror OK" in res.stdout
