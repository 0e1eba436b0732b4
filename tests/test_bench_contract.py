"""CPU checks of bench.py's contract: the reference arm (`--impl reference`, the reference's own OpenMP product on the host cores
through oracle/_ref) prints ONE JSON line with the keys the driver reads and the same `config` dict our arm prints."""
import argparse
import json
import os
import subprocess
import sys

import pytest
from conftest import REPO


def _reference_line(extra, launcher=()):
    cmd = [sys.executable, *launcher, os.path.join(REPO, "bench.py"), "--impl", "reference", "--points", "20000", "--steps", "2", "--warmup", "1", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OPENBLAS_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line(have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref is not built")
    sys.path.insert(0, REPO)
    import bench

    d = _reference_line([])
    assert d["impl"] == "reference" and d["metric"] == "H-matvecs/s" and d["unit"] == "matvec/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0 and abs(d["value"] * d["ms_per_step"] - 1e3) < 1e-6 * 1e3
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same config dict as our arm's for the same arguments
    args = argparse.Namespace(n=20000, mu=1, dtype="double", symmetry="N", gpus=1)
    assert d["config"] == bench.config_dict(args)


def test_reference_arm_under_torchrun_prints_once(have_ref):
    """N > 1: launched like our arm; rank 0 alone runs and prints, the other ranks leave with 0."""
    if not have_ref:
        pytest.skip("oracle/_ref is not built")
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    d = _reference_line(["--gpus", "2"], launcher=("-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port)))
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["config"]["n_partitions"] == 2
