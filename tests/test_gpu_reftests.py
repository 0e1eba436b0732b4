"""The reference's OWN test programs with the GPU twins substituted:

  tests/functional_tests/distributed_operator/test_distributed_operator_product_{double,complex_double}.cpp
  tests/functional_tests/hmatrix/hmatrix_product/test_hmatrix_product_{double,complex_double}.cpp

compiled unmodified from /root/reference by oracle/Makefile (`make reftests`; oracle/ref/reftests/*.cpp explain the
include-time substitution) into oracle/_ref/, which travels to the GPU box prebuilt. Exit code 0 = every check of the
reference's test body (error vs the dense product of the same analytic kernel below the reference's own tolerances) holds
with the H-matrix products running on the B200 through the C ABI.
"""
import os
import subprocess

import pytest
from conftest import REPO

REF_DIR = os.path.join(REPO, "oracle", "_ref")
PROGRAMS = ["reftest_distributed_operator_double", "reftest_distributed_operator_complex_double", "reftest_hmatrix_product_double", "reftest_hmatrix_product_complex_double"]


def _env():
    return dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS=str(min(16, os.cpu_count() or 1)))


@pytest.mark.parametrize("program", PROGRAMS)
def test_twins_are_linked_in(program):
    """(CPU) the substituted programs really call the C ABI: they import the htb_* entry points of libhtool_b200.so."""
    path = os.path.join(REF_DIR, program)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref reference test programs are not built (needs /root/reference)")
    syms = subprocess.run(["nm", "-D", "--undefined-only", path], capture_output=True, text=True, check=True).stdout
    assert "htb_create" in syms and ("htb_add_vector_product" in syms or "htb_add_matrix_product_row_major" in syms)


@pytest.mark.gpu
@pytest.mark.parametrize("program", PROGRAMS)
def test_reference_test_program_passes_on_the_gpu(program):
    path = os.path.join(REF_DIR, program)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref did not travel with the repo")
    r = subprocess.run([path], capture_output=True, text=True, timeout=1500, env=_env())
    tail = (r.stdout[-1500:] + r.stderr[-1500:])
    assert r.returncode == 0, tail
    assert "Errors on a" in r.stdout  # the reference's checks ran and printed their errors
    assert "[htool_b200]" not in r.stdout + r.stderr, tail  # no product was skipped or failed on the device
