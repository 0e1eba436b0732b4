"""CPU tests: pin the plain-C oracle (oracle/hmat_oracle.c) against the reference.

1. against the golden vectors in tests/golden/ (the unmodified reference's sequential_internal_* and
   openmp_internal_* outputs, frozen by tools/make_golden.py);
2. when oracle/_ref is present, against the reference run live on freshly assembled H-matrices.
Tolerance: 1e-13 relative l2 (both sides are FP64 with different summation orders; north_star asks 1e-12).
"""
import numpy as np
import pytest
from conftest import GOLDEN, load_golden, rel_err, rnd, valid_trans

TOL = 1e-13


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_golden(name):
    flat, entries, _ = load_golden(name)
    assert entries
    for e in entries:
        y = e["y_in"].copy()
        if e["mu"] == 1:
            st = flat.oracle_vector_product(e["trans"], e["alpha"], e["x"], e["beta"], y)
        else:
            st = flat.oracle_matrix_product_row_major(e["trans"], e["alpha"], e["x"], e["beta"], y, e["mu"])
        assert st == 0
        assert rel_err(y, e["y_seq"]) < TOL, (name, e["trans"], e["mu"])
        # the reference's two variants only differ by summation order
        assert rel_err(e["y_omp"], e["y_seq"]) < TOL


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_rejects_what_the_reference_rejects(name):
    flat, _, _ = load_golden(name)
    x = np.zeros(max(flat.nb_rows, flat.nb_cols), flat.np_dtype)
    y = np.zeros(max(flat.nb_rows, flat.nb_cols), flat.np_dtype)
    for t in "NTC":
        st = flat.oracle_vector_product(t, 1.0, x, 0.0, y)
        assert (st == 0) == (t in valid_trans(flat.symmetry))


LIVE = [
    dict(n=1500),
    dict(n=1500, symmetry="S", uplo="L"),
    dict(n=1100, dtype="complex", kernel="hermitian_reg", symmetry="H", uplo="U"),
    dict(n=1100, dtype="complex", kernel="complex_reg", symmetry="S", uplo="U"),
    dict(n=900, n_source=700, same_cluster=False, geometry="disk", z_source=0.3, kernel="laplace", epsilon=1e-6),
    dict(n=2000, n_partitions=4, partition_rank=2, symmetry="S", uplo="L"),
    dict(n=1200, dtype="complex", kernel="helmholtz", n_partitions=2, partition_rank=0, symmetry="S", uplo="L"),
]


@pytest.mark.parametrize("kw", LIVE, ids=[str(i) for i in range(len(LIVE))])
def test_oracle_matches_live_reference(kw, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle import refharness as R
    from oracle.flatcase import FlatCase

    case = R.RefCase(**kw)
    flat = FlatCase.from_desc(case.desc)
    info = case.info()
    assert flat.coefficients == info["coefficients"]
    rng = np.random.default_rng(7)
    for trans in valid_trans(info["symmetry_for_leaves"]):
        ni, no = (case.nb_cols, case.nb_rows) if trans == "N" else (case.nb_rows, case.nb_cols)
        alpha, beta = (0.7, -1.3) if case.np_dtype == np.float64 else (0.7 + 0.2j, -1.3 + 0.4j)
        for variant in ("sequential", "openmp"):
            x, y0 = rnd(rng, ni, case.np_dtype), rnd(rng, no, case.np_dtype)
            yr, yo = y0.copy(), y0.copy()
            case.vector_product(trans, alpha, x, beta, yr, variant=variant)
            assert flat.oracle_vector_product(trans, alpha, x, beta, yo) == 0
            assert rel_err(yo, yr) < TOL
        mu = 4
        X, Y0 = rnd(rng, ni * mu, case.np_dtype), rnd(rng, no * mu, case.np_dtype)
        Yr, Yo = Y0.copy(), Y0.copy()
        case.matrix_product_row_major(trans, alpha, X, beta, Yr, mu, variant="openmp")
        assert flat.oracle_matrix_product_row_major(trans, alpha, X, beta, Yo, mu) == 0
        assert rel_err(Yo, Yr) < TOL
