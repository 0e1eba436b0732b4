"""Leaf assembly on the device, second step: the admissible blocks of the block cluster tree are
compressed on the GPU by a batched sympartialACA (htb_create_compressed, htool_b200/csrc/aca.cu) instead of
HMatrixTreeBuilder::openmp_compute_blocks -> HMatrix::compute_low_rank_data -> sympartialACA on the host
(tree_builder.hpp:604-626, hmatrix.hpp:228-237, lrmat/sympartialACA.hpp:41-216).

Given only the block cluster tree, the points and epsilon, the device must rebuild the H-matrix the reference assembled:
the SAME rank for every admissible block and a leaf store BIT-IDENTICAL to the one packed from the reference's factors
(both sides: U panels + dense leaves, V^T panels), hence bit-identical products.
"""
import numpy as np
import pytest
from aca_cases import ACA_GOLDEN, DIAG_FLAGS, AcaCase, packed_side
from conftest import rel_err, rnd

from htool_b200 import capi


def assemble_and_compare(case, mask=None, expected=None, fma_axpy=False, dots=0):
    """Device assembly of `case` with the leaves of `mask` as admissible blocks; `expected` = the FlatCase the store must equal.
    dots: how the teams of 128 / 512 threads compute the stopping criterion (option aca_dots, aca.cuh)."""
    expected = expected or case.flat
    desc0, keep = case.stripped_desc(mask)
    capi.set_option("aca_fma_axpy", 1 if fma_axpy else 0)
    capi.set_option("aca_dots", dots)
    try:
        op = capi.Operator(desc0, generator=(case.kernel, case.tp, case.sp, case.wavenumber), compress_epsilon=case.epsilon)
    finally:
        capi.set_option("aca_fma_axpy", 0)
        capi.set_option("aca_dots", 0)
    exact = case.kernel != "helmholtz"  # (the device's sincos differs from the host's by ulps: same ranks, coefficients to rounding)
    ranks = op.leaf_ranks()
    assert np.array_equal(ranks, expected.table[:, 4]), f"{int((ranks != expected.table[:, 4]).sum())} leaves got another rank than the reference"
    ci = op.compression_info()
    n_blocks = int((expected.table[:, 4] >= 0).sum()) if mask is None else int(np.asarray(mask).sum())
    assert ci["nb_blocks"] == n_blocks
    assert ci["nb_failed"] == n_blocks - int((expected.table[:, 4] >= 0).sum())
    m, n, r = (expected.table[:, k].astype(np.int64) for k in (2, 3, 4))
    assert ci["coefficients"] == int((r * (m + n))[r > 0].sum())
    for side in (0, 1):
        ref_stream, _, _ = packed_side(expected.desc, side, False)
        got = op.download_store(side, ref_stream.size)
        if exact:
            assert np.array_equal(got, ref_stream), f"side {side}: the device-assembled store differs from the one packed from the reference's factors"
        else:  # same layout (same ranks): descriptors identical as bytes, coefficients within rounding of the reference's
            a, b = got.view(np.complex128), ref_stream.view(np.complex128)
            same = (got.view(np.uint64) == ref_stream.view(np.uint64)).reshape(-1, 2).all(axis=1)
            # (absolute, against factor entries of size ~1: the later terms of a cross approximation are differences of such numbers
            # divided by small pivots — one ulp in the kernel function moves entries of rank-11..19 factors by ~3e-11 (measured
            # 1.7e-11 on the fixture), more for the rank-30+ factors of the large blocks; a different PIVOT would move them by O(0.1))
            with np.errstate(invalid="ignore", over="ignore"):
                diff = np.where(same, 0.0, np.abs(a - b))
            # (measured: 1.7e-11 on the fixture, 7e-8 on a few entries of the n = 15000 tree — the cross approximation's pivoting
            # amplifies like Gaussian elimination; the products below, where those errors cancel, agree to 1e-12)
            assert np.isfinite(diff).all() and diff.max() <= 1e-6, f"side {side}: coefficients differ by {diff.max():.2e}: more than rounding"
            assert np.sqrt((diff**2).sum()) <= 1e-10 * np.linalg.norm(expected.coeffs), f"side {side}: {np.sqrt((diff**2).sum()):.2e}"
    rng = np.random.default_rng(11)
    dt = case.dtype
    sym = case.flat.symmetry
    op_ref = capi.Operator(expected.desc)
    for trans in ("N", "C" if sym == "H" else "T"):
        ni, no = (case.flat.nb_cols, case.flat.nb_rows) if trans == "N" else (case.flat.nb_rows, case.flat.nb_cols)
        x = rnd(rng, ni, dt)
        y, y_ref, y_or = np.zeros(no, dt), np.zeros(no, dt), np.zeros(no, dt)
        op.add_vector_product(trans, 1.0, x, 0.0, y)
        op_ref.add_vector_product(trans, 1.0, x, 0.0, y_ref)
        expected.oracle_vector_product(trans, 1.0, x, 0.0, y_or)
        assert np.array_equal(y, y_ref) if exact else rel_err(y, y_ref) < 1e-12
        assert rel_err(y, y_or) < 1e-12
    mu = 8
    X = rnd(rng, case.flat.nb_cols * mu, dt)
    Y, Y_ref = np.zeros(case.flat.nb_rows * mu, dt), np.zeros(case.flat.nb_rows * mu, dt)
    op.add_matrix_product_row_major("N", 1.0, X, 0.0, Y, mu)
    op_ref.add_matrix_product_row_major("N", 1.0, X, 0.0, Y_ref, mu)
    assert np.array_equal(Y, Y_ref) if exact else rel_err(Y, Y_ref) < 1e-12
    op_ref.close()
    return op


@pytest.mark.gpu
@pytest.mark.parametrize("name", ACA_GOLDEN)
@pytest.mark.parametrize("dots", [0, 1, 2], ids=["warp_dots_guarded", "in_order_dots", "replay_always"])
def test_device_assembly_rebuilds_the_reference_hmatrix(name, dots):
    case = AcaCase.golden(name)
    op = assemble_and_compare(case, dots=dots)
    y = np.zeros(case.flat.nb_rows, case.dtype)
    op.add_vector_product("N", 1.0, case.x, 0.0, y)
    assert rel_err(y, case.y) < 1e-12  # the reference's own product on its own H-matrix
    op.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["d_N", "d_SL", "d_rect"])
@pytest.mark.parametrize("fma_axpy", [False, True], ids=["mul_add", "fma"])
@pytest.mark.parametrize("epsilon", [None, 0.05], ids=["eps_fixture", "eps_0.05"])
def test_every_offdiagonal_leaf_as_an_admissible_block(name, fma_axpy, epsilon):
    """Blocks the reference never compresses (its near-field leaves: small and hard; at the fixture's epsilon they all FAIL and
    stay dense — the reference's "false positives", tree_builder.hpp:619-625 —, at epsilon = 0.05 most of them compress)
    against the oracle's sympartialACA, for both BLAS axpy variants."""
    case = AcaCase.golden(name)
    if epsilon:
        case.epsilon = epsilon
    f = case.flat
    mask = (f.table[:, 5] & DIAG_FLAGS) == 0
    if f.symmetry == "N" and f.nb_rows == f.nb_cols:
        mask &= f.table[:, 0] != f.table[:, 1]
    expected = case.reassembled(mask, fma_axpy)
    n_failed = int(((expected.table[:, 4] < 0) & mask).sum())
    assert n_failed > 0 and (not epsilon or int((expected.table[:, 4] > 0).sum()) > int((f.table[:, 4] > 0).sum()))
    assemble_and_compare(case, mask, expected, fma_axpy).close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(n=20000, kernel="laplace_reg", epsilon=1e-4), dict(n=12000, kernel="laplace_reg", epsilon=1e-6, symmetry="S", uplo="L"),
                                dict(n=9000, n_partitions=4, partition_rank=2, kernel="laplace_reg", epsilon=1e-4),
                                dict(n=12000, dtype="complex", kernel="complex_reg", epsilon=1e-4, symmetry="S", uplo="L"),
                                dict(n=10000, dtype="complex", kernel="hermitian_reg", epsilon=1e-4, symmetry="H", uplo="U"),
                                dict(n=15000, dtype="complex", kernel="helmholtz", epsilon=1e-4, wavenumber=5.0),
                                dict(n=8000, dtype="complex", kernel="helmholtz", epsilon=1e-4, wavenumber=5.0, symmetry="S", uplo="L", n_partitions=2, partition_rank=1)],
                         ids=["N20000", "SL12000_eps6", "strip", "z_SL12000", "z_HU10000", "z_helmholtz15000", "z_helmholtz_S_strip"])
def test_device_assembly_against_the_live_reference(kw, have_ref):
    """Larger trees (blocks of all three team sizes, ranks up to ~30), assembled by the reference on this machine's cores."""
    if not have_ref:
        pytest.skip("oracle/_ref is not built")
    case = AcaCase.live(**kw)
    for dots in (0, 1, 2):  # (the guarded warp-level dot products, the reference's in-order ones, the in-order replay at every iteration)
        assemble_and_compare(case, dots=dots).close()


@pytest.mark.gpu
def test_pool_overflow_is_retried():
    """A factor pool sized for one term per block overflows; the batch is repeated with a larger pool and gives the same store."""
    case = AcaCase.golden("d_SL")
    capi.set_option("aca_rank_guess", 1)
    try:
        assemble_and_compare(case).close()
    finally:
        capi.set_option("aca_rank_guess", 16)


@pytest.mark.gpu
def test_compress_arguments_are_validated():
    case = AcaCase.golden("d_N")
    desc0, keep = case.stripped_desc()
    with pytest.raises(capi.HtbError) as ei:  # a complex kernel function cannot fill a double H-matrix
        capi.Operator(desc0, generator=("helmholtz", case.tp, case.sp, 1.0), compress_epsilon=1e-4)
    assert ei.value.status == capi.HTB_ERR_INVALID
    with pytest.raises(capi.HtbError) as ei:
        capi.Operator(desc0, generator=(case.kernel, case.tp, case.sp, 0.0), compress_epsilon=0.0)
    assert ei.value.status == capi.HTB_ERR_INVALID
    with pytest.raises(capi.HtbError) as ei:  # plain htb_create_generated refuses blocks that are still to compress
        capi.Operator(desc0, generator=(case.kernel, case.tp, case.sp, 0.0))
    assert ei.value.status == capi.HTB_ERR_INVALID
    lv = np.frombuffer(keep, dtype=capi.LEAF_NP_DTYPE)
    i = int(np.nonzero(lv["rank"] == capi.HTB_RANK_COMPRESS)[0][0])
    lv["data0"][i] = case.flat.coeffs.ctypes.data
    with pytest.raises(capi.HtbError) as ei:  # a block to compress cannot carry data
        capi.Operator(desc0, generator=(case.kernel, case.tp, case.sp, 0.0), compress_epsilon=1e-4)
    assert ei.value.status == capi.HTB_ERR_INVALID
