"""Drop-in tests (-m gpu): the reference's own operator / DistributedOperator code with the GPU twins of
htool_b200/cpp/htool_b200/operators.hpp plugged in, against the same code with the reference's CPU adapters,
on the same HMatrix object. The C++ side is oracle/ref/dropin_capi.cpp, compiled against the
unmodified reference headers into oracle/_ref/libhtool_dropin.so (prebuilt in the build container, it travels
to the GPU box with the repository snapshot).

Every group is the worst relative l2 difference between the two runs; tolerance 1e-12 (BASELINE.json).
"""
import ctypes as C
import os

import numpy as np
import pytest
from conftest import REPO

pytestmark = pytest.mark.gpu

DROPIN_LIB = os.path.join(REPO, "oracle", "_ref", "libhtool_dropin.so")
GROUPS = ["g2l_vector", "g2l_row_major", "g2l_sub_product", "l2l_vector", "l2l_row_major", "l2l_sub_product", "dist_vector_g2g",
          "dist_vector_l2l", "dist_matrix_g2g", "dist_matrix_l2l", "dist_row_major_l2l", "dist_sub_product", "free_vector_user",
          "free_matrix_user", "logged_unsupported", "device_dist", "generated_dense", "device_assembly"]

CASES = [
    dict(n=6000),
    dict(n=6000, symmetry="S", uplo="L"),
    dict(n=5000, symmetry="S", uplo="U"),
    dict(n=4000, dtype="complex", kernel="complex_reg", symmetry="S", uplo="L"),
    dict(n=4000, dtype="complex", kernel="hermitian_reg", symmetry="H", uplo="L"),
    dict(n=4000, dtype="complex", kernel="helmholtz"),
    dict(n=5000, n_source=3500, same_cluster=False, geometry="disk", z_source=0.4, kernel="laplace", epsilon=1e-6),
    dict(n=8000, n_partitions=4, partition_rank=2),                                  # a row strip: operator-level checks only
    dict(n=8000, n_partitions=4, partition_rank=1, symmetry="S", uplo="L"),
    dict(n=8000, n_partitions=2, partition_rank=1, local_block=True, symmetry="S", uplo="L"),  # diagonal block, local-to-local
]


@pytest.mark.parametrize("kw", CASES, ids=[str(i) for i in range(len(CASES))])
def test_reference_code_runs_unchanged_on_the_gpu_operators(kw):
    if not os.path.exists(DROPIN_LIB):
        pytest.skip("oracle/_ref/libhtool_dropin.so did not travel with the repo")
    import torch

    assert torch.cuda.is_available()
    from oracle import refharness as R

    R.load()  # pins OPENBLAS_NUM_THREADS=1 before any BLAS call from inside an OpenMP region
    lib = C.CDLL(DROPIN_LIB)
    lib.dropin_n_groups.restype = C.c_int
    lib.dropin_run.restype = C.c_int
    lib.dropin_run.argtypes = [C.POINTER(R.ref_case_spec), C.c_void_p, C.c_int]
    assert lib.dropin_n_groups() == len(GROUPS)
    spec = R.make_spec(**kw)
    out = np.full(len(GROUPS), -1.0)
    rc = lib.dropin_run(C.byref(spec), out.ctypes.data, out.size)
    assert rc == 0, f"dropin_run failed with {rc}"
    whole = kw.get("partition_rank", -1) < 0
    for name, err in zip(GROUPS, out):
        if err < 0:  # group not run for this case: distributed / free-function groups need the whole operator
            assert (name.startswith(("dist_", "free_", "device_dist")) and not whole) or (name == "logged_unsupported" and kw.get("symmetry", "N") == "N") \
                or (name == "device_dist" and not kw.get("same_cluster", True)) or (name == "generated_dense" and kw.get("compressor", "sympartialACA") != "sympartialACA") \
                or (name == "device_assembly" and kw.get("compressor", "sympartialACA") != "sympartialACA"), name
            continue
        assert err < 1e-12, (kw, name, err)
    # the operator-level groups always run
    assert all(out[i] >= 0 for i in range(6))
