"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI
(htool_b200/lib/libhtool_b200.so via ctypes, the same entry points the C++ shim binds), against

  * the golden vectors produced by the unmodified reference (tests/golden/, tools/make_golden.py),
  * the plain-C oracle on seeded random leaf lists (ragged / overlapping / rank-0 / 1x1 shapes),
  * the reference itself run live through oracle/_ref when that prebuilt library travelled with the repo.

Tolerance: relative l2 error <= 1e-12 (BASELINE.json north_star) for double and complex<double>.
"""
import numpy as np
import pytest
from conftest import GOLDEN, load_golden, rel_err, rnd, valid_trans

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def capi():
    import torch

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from htool_b200 import capi as m

    m.load()
    return m


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_vectors(capi, name):
    flat, entries, z = load_golden(name)
    op = capi.Operator(flat.desc)
    info = op.info()
    assert info["coefficients"] == flat.coefficients
    assert info["coefficients_twice"] == flat.coefficients_twice
    before = op.launch_count()
    for e in entries:
        y = e["y_in"].copy()
        if e["mu"] == 1:
            op.add_vector_product(e["trans"], e["alpha"], e["x"], e["beta"], y)
        else:
            op.add_matrix_product_row_major(e["trans"], e["alpha"], e["x"], e["beta"], y, e["mu"])
        assert rel_err(y, e["y_seq"]) < TOL, (name, e["trans"], e["mu"], rel_err(y, e["y_seq"]))
        assert rel_err(y, e["y_omp"]) < TOL
    assert op.launch_count() > before  # the CUDA kernels ran (no fallback exists)
    # rejected combinations: same condition as add_hmatrix_vector_product.hpp:112-115
    for t in "NTC":
        if t not in valid_trans(flat.symmetry):
            with pytest.raises(capi.HtbError) as ei:
                op.add_vector_product(t, 1.0, np.zeros(flat.nb_rows, flat.np_dtype), 0.0, np.zeros(flat.nb_cols, flat.np_dtype))
            assert ei.value.status == capi.HTB_ERR_UNSUPPORTED
    if "user_x" in z:
        op.set_permutations(z["perm_target"], z["perm_source"])
        a, b = z["user_ab"]
        y = z["user_yin"].copy()
        op.add_vector_product_user_numbering("N", a, z["user_x"], b, y)
        assert rel_err(y, z["user_yref"]) < TOL
        Y = z["userm_yin"].copy()
        op.add_matrix_product_user_numbering("N", a, z["userm_x"], b, Y, 2)
        assert rel_err(Y, z["userm_yref"]) < TOL
    op.close()


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("dtype_code,symmetric", [(0, None), (1, None), (0, "S"), (1, "S"), (1, "H")])
def test_random_leaf_lists_vs_oracle(capi, seed, dtype_code, symmetric):
    from oracle.flatcase import random_flatcase

    flat = random_flatcase(seed=seed, dtype_code=dtype_code, symmetric=symmetric, nb_rows=700, nb_cols=530, n_leaves=150, max_dim=300, max_rank=40)
    op = capi.Operator(flat.desc)
    rng = np.random.default_rng(seed)
    for trans in valid_trans(flat.symmetry):
        ni, no = (flat.nb_cols, flat.nb_rows) if trans == "N" else (flat.nb_rows, flat.nb_cols)
        for alpha, beta in [(1.0, 0.0), (0.7, -1.3)]:
            if flat.np_dtype == np.complex128:
                alpha, beta = alpha * (1 + 0.5j), beta * (1 - 0.25j)
            x, y0 = rnd(rng, ni, flat.np_dtype), rnd(rng, no, flat.np_dtype)
            yo, yg = y0.copy(), y0.copy()
            assert flat.oracle_vector_product(trans, alpha, x, beta, yo) == 0
            op.add_vector_product(trans, alpha, x, beta, yg)
            assert rel_err(yg, yo) < TOL, (trans, alpha, beta, rel_err(yg, yo))
        # mu >= 2: FP64 tensor-core kernels on RUNS (mkernels.cu), column groups of 64 real / 32 complex right-hand sides
        am, bm = (0.5, 2.0) if flat.np_dtype == np.float64 else (0.5 - 0.3j, 2.0 + 0.25j)
        for mu in (2, 5, 8, 13, 24, 33, 64, 70):
            X, Y0 = rnd(rng, ni * mu, flat.np_dtype), rnd(rng, no * mu, flat.np_dtype)
            Yo, Yg = Y0.copy(), Y0.copy()
            flat.oracle_matrix_product_row_major(trans, am, X, bm, Yo, mu)
            op.add_matrix_product_row_major(trans, am, X, bm, Yg, mu)
            assert rel_err(Yg, Yo) < TOL, (trans, mu, rel_err(Yg, Yo))
            if mu == 13:  # beta == 0 must not read the output
                Yg = np.full(no * mu, np.nan, flat.np_dtype)
                Yo = np.zeros(no * mu, flat.np_dtype)
                flat.oracle_matrix_product_row_major(trans, am, X, 0.0, Yo, mu)
                op.add_matrix_product_row_major(trans, am, X, 0.0, Yg, mu)
                assert rel_err(Yg, Yo) < TOL
    op.close()


@pytest.mark.parametrize("dtype_code", [0, 1])
def test_tensor_core_path_from_two_right_hand_sides(capi, dtype_code):
    """mrhs_min = 2 forces the DMMA kernels for mu = 2, 3, 4 (by default a loop of single-RHS products is faster there)."""
    from oracle.flatcase import random_flatcase

    flat = random_flatcase(seed=11, dtype_code=dtype_code, symmetric="S", nb_rows=500, nb_cols=500, n_leaves=120, max_dim=200, max_rank=30)
    capi.set_option("mrhs_min", 2)
    try:
        op = capi.Operator(flat.desc)
        rng = np.random.default_rng(2)
        for trans in valid_trans(flat.symmetry):
            for mu in (2, 3, 4):
                X, Y0 = rnd(rng, flat.nb_cols * mu, flat.np_dtype), rnd(rng, flat.nb_rows * mu, flat.np_dtype)
                Yo, Yg = Y0.copy(), Y0.copy()
                l0 = op.launch_count()
                flat.oracle_matrix_product_row_major(trans, 0.5, X, -1.5, Yo, mu)
                op.add_matrix_product_row_major(trans, 0.5, X, -1.5, Yg, mu)
                assert rel_err(Yg, Yo) < TOL, (trans, mu)
                # one multi-RHS pass sequence (staging of the group + twice under symmetry + the near-field pass and, the first time, the
                # two launches that build the near-field panels), not mu products of three passes each
                assert op.launch_count() - l0 <= 10
        op.close()
    finally:
        capi.set_option("mrhs_min", 0)


def test_beta_zero_ignores_nan_in_out(capi):
    flat, entries, _ = load_golden("d_N")
    op = capi.Operator(flat.desc)
    e = entries[0]
    y = np.full(flat.nb_rows, np.nan)
    op.add_vector_product("N", 1.0, e["x"], 0.0, y)
    ref = np.zeros(flat.nb_rows)
    flat.oracle_vector_product("N", 1.0, e["x"], 0.0, ref)
    assert rel_err(y, ref) < TOL
    op.close()


def test_deterministic_bitwise(capi):
    flat, entries, _ = load_golden("d_SL")
    op = capi.Operator(flat.desc)
    e = entries[1]
    outs = []
    for _ in range(5):
        y = e["y_in"].copy()
        op.add_vector_product(e["trans"], e["alpha"], e["x"], e["beta"], y)
        outs.append(y)
    for y in outs[1:]:
        assert np.array_equal(y, outs[0])  # fixed summation order: bit-identical across runs
    op.close()


@pytest.mark.parametrize("opts", [dict(block_rows=32, piece_cols=4, stage_bytes=4096, cseg_bytes=512, ring_stages=2, reduce_ring_stages=9), dict(block_rows=128, piece_cols=16, stage_bytes=16448, cseg_bytes=2048),
                                  dict(evict_first=0, ring_stages=5, reduce_ring_stages=2, stage_bytes=8192, piece_cols=8, cseg_bytes=2048),
                                  dict(fused_symmetric=0)])
def test_packer_and_launch_options(capi, opts):
    defaults = {k: capi.get_option(k) for k in ("block_rows", "piece_cols", "stage_bytes", "cseg_bytes", "ring_stages", "reduce_ring_stages", "evict_first", "fused_symmetric")}
    try:
        for k, v in opts.items():
            capi.set_option(k, v)
        for name in ("d_SL", "z_HU", "d_strip_SU"):
            flat, entries, _ = load_golden(name)
            if flat.np_dtype == np.complex128 and opts.get("block_rows", 64) > 64:
                with pytest.raises(capi.HtbError):  # block_rows * sizeof(T) <= 1024
                    capi.Operator(flat.desc)
                continue
            op = capi.Operator(flat.desc)
            for e in entries:
                if e["mu"] != 1:
                    continue
                y = e["y_in"].copy()
                op.add_vector_product(e["trans"], e["alpha"], e["x"], e["beta"], y)
                assert rel_err(y, e["y_seq"]) < TOL
            op.close()
    finally:
        for k, v in defaults.items():
            capi.set_option(k, v)


def test_device_pointers_and_caller_stream(capi):
    import torch

    flat, entries, _ = load_golden("d_N")
    op = capi.Operator(flat.desc)
    e = entries[1]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        op.set_stream(stream.cuda_stream)
        x = torch.from_numpy(e["x"]).cuda()
        y = torch.from_numpy(e["y_in"].copy()).cuda()
        op.add_vector_product_device(e["trans"], e["alpha"], x.data_ptr(), e["beta"], y.data_ptr())
        stream.synchronize()
        assert rel_err(y.cpu().numpy(), e["y_seq"]) < TOL
    op.set_stream(None)
    op.close()


def test_empty_operator_scales_out(capi):
    from oracle.flatcase import FlatCase

    f = FlatCase(0, 37, 21, 0, 0, "N", "N", np.zeros((0, 6), np.int32), np.zeros(0))
    op = capi.Operator(f.desc)
    x, y = np.ones(21), np.full(37, 2.0)
    op.add_vector_product("N", 1.0, x, 0.5, y)
    assert np.allclose(y, 1.0)
    op.close()


LIVE = [
    dict(n=20000),
    dict(n=20000, symmetry="S", uplo="L"),
    dict(n=12000, dtype="complex", kernel="helmholtz", symmetry="S", uplo="L"),
    dict(n=12000, dtype="complex", kernel="hermitian_reg", symmetry="H", uplo="U"),
    dict(n=30000, n_partitions=4, partition_rank=1, symmetry="S", uplo="L"),
    dict(n=9000, n_source=6000, same_cluster=False, geometry="disk", z_source=0.3, kernel="laplace", epsilon=1e-6),
]


@pytest.mark.parametrize("kw", LIVE, ids=[str(i) for i in range(len(LIVE))])
def test_live_reference(capi, kw, have_ref):
    """Same compressed HMatrix object -> reference openmp_internal_* vs GPU."""
    if not have_ref:
        pytest.skip("oracle/_ref did not travel with the repo")
    from oracle import refharness as R

    case = R.RefCase(**kw)
    info = case.info()
    op = capi.Operator(case.desc)
    assert op.info()["coefficients"] == info["coefficients"]
    rng = np.random.default_rng(11)
    for trans in valid_trans(info["symmetry_for_leaves"]):
        ni, no = (case.nb_cols, case.nb_rows) if trans == "N" else (case.nb_rows, case.nb_cols)
        alpha, beta = (0.7, -1.3) if case.np_dtype == np.float64 else (0.7 + 0.2j, -1.3 + 0.4j)
        x, y0 = rnd(rng, ni, case.np_dtype), rnd(rng, no, case.np_dtype)
        yr, yg = y0.copy(), y0.copy()
        case.vector_product(trans, alpha, x, beta, yr, variant="openmp")
        op.add_vector_product(trans, alpha, x, beta, yg)
        assert rel_err(yg, yr) < TOL, (kw, trans, rel_err(yg, yr))
    for mu in (3, 16):
        X, Y0 = rnd(rng, case.nb_cols * mu, case.np_dtype), rnd(rng, case.nb_rows * mu, case.np_dtype)
        Yr, Yg = Y0.copy(), Y0.copy()
        case.matrix_product_row_major("N", 1.0, X, 0.5, Yr, mu, variant="openmp")
        op.add_matrix_product_row_major("N", 1.0, X, 0.5, Yg, mu)
        assert rel_err(Yg, Yr) < TOL, (kw, mu, rel_err(Yg, Yr))
    # size-independent property: linearity in x
    x1, x2 = rnd(rng, case.nb_cols, case.np_dtype), rnd(rng, case.nb_cols, case.np_dtype)
    y1, y2, y12 = (np.zeros(case.nb_rows, case.np_dtype) for _ in range(3))
    op.add_vector_product("N", 1.0, x1, 0.0, y1)
    op.add_vector_product("N", 1.0, x2, 0.0, y2)
    op.add_vector_product("N", 1.0, 2.0 * x1 - 3.0 * x2, 0.0, y12)
    assert rel_err(y12, 2.0 * y1 - 3.0 * y2) < 1e-11
    op.close()


def test_full_size_properties(capi, have_ref):
    """BASELINE.json configs[1] at FULL size (Laplace N = 1e6, eps = 1e-4): parity against the reference's OpenMP product
    on the same compressed HMatrix, plus size-independent properties of the GPU path alone — linearity in x, the
    beta term, the adjoint identity <H x, y> = <x, H^T y> through the transposed passes, bit-reproducibility, and the
    device-pointer path agreeing with the host-pointer path."""
    if not have_ref:
        pytest.skip("oracle/_ref did not travel with the repo")
    import torch

    import bench
    from oracle import refharness as R

    R.set_num_threads(__import__("os").cpu_count() or 1)
    case = R.RefCase(**bench.case_kwargs(1_000_000, "double", "N"))
    n = case.nb_rows
    op = capi.Operator(case.desc)
    assert op.info()["coefficients"] == case.info()["coefficients"] > 2_000_000_000
    rng = np.random.default_rng(17)
    x1, x2, w = rng.random(n) - 0.5, rng.random(n) - 0.5, rng.random(n) - 0.5
    y1, y2, y12, yr = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    op.add_vector_product("N", 1.0, x1, 0.0, y1)
    case.vector_product("N", 1.0, x1, 0.0, yr, variant="openmp")
    assert rel_err(y1, yr) < TOL
    op.add_vector_product("N", 1.0, x2, 0.0, y2)
    op.add_vector_product("N", 1.0, 2.0 * x1 - 3.0 * x2, 0.0, y12)
    assert rel_err(y12, 2.0 * y1 - 3.0 * y2) < 1e-11  # linearity
    yb = w.copy()
    op.add_vector_product("N", 0.5, x1, -2.0, yb)
    assert rel_err(yb, 0.5 * y1 - 2.0 * w) < 1e-11  # alpha / beta
    z = np.zeros(n)
    op.add_vector_product("T", 1.0, w, 0.0, z)
    lhs, rhs = float(y1 @ w), float(x1 @ z)
    assert abs(lhs - rhs) <= 1e-10 * (np.linalg.norm(y1) * np.linalg.norm(w))  # adjoint identity: REDUCE/APPLY roles swapped
    again = np.zeros(n)
    op.add_vector_product("N", 1.0, x1, 0.0, again)
    assert np.array_equal(again, y1)  # fixed summation order
    xd, yd = torch.from_numpy(x1).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda")
    op.add_vector_product_device("N", 1.0, xd.data_ptr(), 0.0, yd.data_ptr())
    op.synchronize()
    assert np.array_equal(yd.cpu().numpy(), y1)
    # multi-RHS at full size on the tensor-core path: column c of the result equals the single-RHS product of column c
    mu = 8
    X = np.ascontiguousarray(np.stack([x1, x2, w, x1 + x2, x1 - w, 2 * x2, -x1, w + x2], axis=1))
    Y = np.zeros((n, mu))
    op.add_matrix_product_row_major("N", 1.0, X.reshape(-1), 0.0, Y.reshape(-1), mu)
    assert rel_err(Y[:, 0], y1) < TOL and rel_err(Y[:, 1], y2) < TOL
    assert rel_err(Y[:, 3], y1 + y2) < 1e-11
    op.close()


@pytest.mark.parametrize("name", ["d_N", "d_SL", "z_HU"])
def test_page_locked_host_vectors_zero_copy(capi, name):
    """htb_host_register: the kernels then read x from / write y to the caller's HOST buffers directly (zero copy, mu = 1),
    or DMA them without staging (zero_copy = 0, mu > 1). Same bits as the pageable path."""
    flat, entries, _ = load_golden(name)
    op = capi.Operator(flat.desc)
    for e in entries:
        y_page = e["y_in"].copy()
        x_pin, y_pin = e["x"].copy(), e["y_in"].copy()

        def run(x, y):
            if e["mu"] == 1:
                op.add_vector_product(e["trans"], e["alpha"], x, e["beta"], y)
            else:
                op.add_matrix_product_row_major(e["trans"], e["alpha"], x, e["beta"], y, e["mu"])

        run(e["x"], y_page)
        capi.host_register(x_pin)
        capi.host_register(y_pin)
        try:
            run(x_pin, y_pin)
            assert np.array_equal(y_pin, y_page)
            capi.set_option("zero_copy", 0)
            y_pin[:] = e["y_in"]
            run(x_pin, y_pin)
            assert np.array_equal(y_pin, y_page)
        finally:
            capi.set_option("zero_copy", 1)
            capi.host_unregister(x_pin)
            capi.host_unregister(y_pin)
        assert rel_err(y_page, e["y_seq"]) < TOL
    op.close()


@pytest.mark.parametrize("nf_rows", [0, 64, 24])
@pytest.mark.parametrize("dtype_code,symmetric", [(0, None), (1, None), (0, "S"), (1, "H")])
def test_near_field_panels_of_the_multi_rhs_product(capi, dtype_code, symmetric, nf_rows):
    """Option m_near_field = 1: for the multi-RHS product 'N' the dense leaves of a target block are applied from ONE
    block-sparse panel per block (or per group of m_nf_rows rows), built on the device from the main stream, and skipped in
    the runs of the main stream (RunDesc::K_lr). Same results — on a block tree, and on overlapping leaf lists where the
    panels must add up."""
    from oracle.flatcase import random_flatcase

    capi.set_option("m_near_field", 1)
    capi.set_option("m_nf_rows", nf_rows)
    try:
        cases = [random_flatcase(seed=3, dtype_code=dtype_code, symmetric=symmetric, nb_rows=700, nb_cols=530, n_leaves=150, max_dim=300, max_rank=40)]
        if symmetric is None:
            cases.append(load_golden("d_N" if dtype_code == 0 else "z_N_helmholtz")[0])
        for flat in cases:
            op = capi.Operator(flat.desc)
            rng = np.random.default_rng(5)
            for trans in valid_trans(flat.symmetry):
                ni, no = (flat.nb_cols, flat.nb_rows) if trans == "N" else (flat.nb_rows, flat.nb_cols)
                for mu in (8, 33, 70):
                    am, bm = (0.5, 2.0) if flat.np_dtype == np.float64 else (0.5 - 0.3j, 2.0 + 0.25j)
                    X, Y0 = rnd(rng, ni * mu, flat.np_dtype), rnd(rng, no * mu, flat.np_dtype)
                    Yo, Yg = Y0.copy(), Y0.copy()
                    flat.oracle_matrix_product_row_major(trans, am, X, bm, Yo, mu)
                    op.add_matrix_product_row_major(trans, am, X, bm, Yg, mu)
                    assert rel_err(Yg, Yo) < TOL, (trans, mu, rel_err(Yg, Yo))
            op.close()
    finally:
        capi.set_option("m_near_field", 0)
        capi.set_option("m_nf_rows", 0)
