"""The user-facing C++ example (examples/use_b200_operator.cpp) compiles against the UNMODIFIED reference headers, and
without a CUDA device every GPU entry point fails loudly through Htool's logger — no CPU fallback computes in its place
(the outputs stay untouched: relative error 1 against the reference adapter)."""
import os
import subprocess

import pytest
from conftest import REPO

EXAMPLE = os.path.join(REPO, "oracle", "_ref", "use_b200_operator")


def test_example_builds_and_fails_loudly_without_a_gpu():
    if not os.path.isdir("/root/reference/include/htool"):
        pytest.skip("the reference headers are only present in the build container")
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the loud-failure path cannot be observed")
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "example"], check=True, stdout=subprocess.DEVNULL)
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="2")
    r = subprocess.run([EXAMPLE, "3000"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout + r.stderr
    assert "htool_b200 has no CPU fallback" in out
    assert "DistributedOperator, GPU twin vs reference adapter: 1" in out  # y untouched: nothing was computed on the CPU instead
    assert "device-resident GMRES: 0 iterations" in out


@pytest.mark.gpu
def test_example_runs_on_the_gpu():
    """The same program on a B200: every way of using the device product — and the leaf assembly on the device — agrees with
    the reference's CPU operator on the same problem."""
    import re

    if not os.path.exists(EXAMPLE):
        pytest.skip("oracle/_ref/use_b200_operator did not travel with the repo")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    r = subprocess.run([EXAMPLE, "20000"], capture_output=True, text=True, timeout=600, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    errors = {}
    for line in out.splitlines():
        m = re.match(r"^(.*?):\s+([0-9.eE+-]+)$", line)
        if m:
            errors[m.group(1).strip()] = float(m.group(2))
    for key in ("DistributedOperator, GPU twin vs reference adapter", "DeviceDistributedOperator vs reference adapter", "add_hmatrix_vector_product (user numbering)",
                "device-assembled H-matrix vs host-assembled (strip product)", "DeviceDistributedOperator over device-assembled strips"):
        assert key in errors, (key, out[-2000:])
        assert errors[key] < 1e-12, (key, errors[key])
    m = re.search(r"device-resident GMRES: (\d+) iterations, relative residual ([0-9.eE+-]+)", out)
    assert m and int(m.group(1)) > 0 and float(m.group(2)) < 1e-6, out[-2000:]
