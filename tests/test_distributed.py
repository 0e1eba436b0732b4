"""Row-sharded (DistributedOperator-style) product, world_size 2.

CPU (`-m "not gpu"`): two gloo ranks; each assembles its row strip with the reference, gathers x with
torch.distributed and runs the strip through the stream emulator — host-side sharding logic, no device.
GPU (`-m gpu`, needs >= 2 GPUs): two NCCL ranks through htb_comm_init / htb_dist_add_product_local_to_local.
"""
import os
import socket
import subprocess
import sys

import pytest
from conftest import REPO

WORKER = os.path.join(REPO, "tests", "dist_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(backend, world, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           WORKER, "--backend", backend] + extra
    env = dict(os.environ, OMP_NUM_THREADS="2", OPENBLAS_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "worst rel. l2 error" in r.stdout
    return r.stdout


@pytest.mark.parametrize("extra", [["--points", "3000"], ["--points", "2500", "--sym", "S"], ["--points", "2000", "--scalar", "complex", "--sym", "S"]], ids=["double", "double_S", "complex_S"])
def test_two_gloo_ranks_host_logic(extra, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref is not built")
    _run("gloo", 2, extra)


def test_four_gloo_ranks_host_logic(have_ref):
    """Four row strips (unequal partition sizes, symmetric storage inside the diagonal blocks only)."""
    if not have_ref:
        pytest.skip("oracle/_ref is not built")
    _run("gloo", 4, ["--points", "3500", "--sym", "S"])


def test_eight_gloo_ranks_host_logic(have_ref):
    """Eight row strips, the partitioning of the 8-GPU scaling run (complex, symmetric storage inside the diagonal blocks)."""
    if not have_ref:
        pytest.skip("oracle/_ref is not built")
    _run("gloo", 8, ["--points", "6000", "--scalar", "complex", "--sym", "S"])


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [["--points", "20000"], ["--points", "16000", "--sym", "S"], ["--points", "12000", "--scalar", "complex", "--sym", "S"], ["--points", "8000", "--rhs", "3"],
                                   ["--points", "8000", "--rhs", "16"], ["--points", "16000", "--sym", "S", "--p2p", "0"], ["--points", "8000", "--rhs", "3", "--p2p", "0"]],
                         ids=["double", "double_S", "complex_S", "mu3", "mu16_dmma", "double_S_nccl_gather", "mu3_nccl_gather"])
def test_two_nccl_ranks(extra, have_ref):
    import torch

    if not have_ref:
        pytest.skip("oracle/_ref did not travel with the repo")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run("nccl", 2, extra)
