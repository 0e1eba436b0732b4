import glob
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
GOLDEN = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build the checkers if they are missing (cheap; the product library is built by __graft_entry__.build)
    import __graft_entry__ as g

    g.build_oracle(quiet=True)
    try:
        g.build_product(quiet=True)
    except (OSError, subprocess.CalledProcessError) as ex:  # no nvcc on this machine: only the tests that load the library fail
        print(f"[conftest] product library not rebuilt ({ex}); tests that need libhtool_b200.so will fail or skip", file=sys.stderr)


def load_golden(name):
    from oracle.flatcase import FlatCase

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    flat = FlatCase.from_arrays(z)
    entries = []
    for k in range(int(z["n_entries"][0])):
        trans, mu = chr(int(z[f"e{k}_meta"][0])), int(z[f"e{k}_meta"][1])
        alpha, beta = z[f"e{k}_ab"]
        entries.append(dict(trans=trans, mu=mu, alpha=alpha, beta=beta, x=z[f"e{k}_x"], y_in=z[f"e{k}_yin"], y_seq=z[f"e{k}_yseq"], y_omp=z[f"e{k}_yomp"]))
    return flat, entries, z


def rel_err(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def rnd(rng, n, dt):
    v = rng.random(n) - 0.5
    if dt == np.complex128:
        v = v + 1j * (rng.random(n) - 0.5)
    return v.astype(dt)


def valid_trans(symmetry):
    return [t for t in "NTC" if not ((t == "T" and symmetry == "H") or (t == "C" and symmetry == "S"))]


@pytest.fixture(scope="session")
def have_ref():
    from oracle import refharness

    return refharness.available()
