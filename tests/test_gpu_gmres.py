"""GPU tests of the device-resident GMRES (htb_gmres, htool_b200/csrc/gmres.cu) against oracle/gmres_oracle.py.

The operator inside both is the same compressed H-matrix: the GPU solver multiplies with the CUDA product, the oracle
with the plain-C oracle product (oracle/hmat_oracle.c) on the same leaves. The Krylov arithmetic of the reference lives
in HPDDM (not vendored: parity unpinned, see oracle/gmres_oracle.py), so the checks are: same iteration count and
matching iterates for a fixed number of iterations, convergence to the requested tolerance measured with the TRUE
residual, and agreement of the converged solutions.
"""
import numpy as np
import pytest
from conftest import load_golden, rel_err, rnd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    import torch

    assert torch.cuda.is_available()
    from htool_b200 import capi as m

    m.load()
    return m


def _shifted(flat, shift):
    """matvec of (H + shift I): the golden operators are first-kind kernels, the shift makes the test systems well posed
    without changing what is exercised (the H-matrix product)."""
    def mv(v):
        y = shift * v
        flat.oracle_vector_product("N", 1.0, np.ascontiguousarray(v), 1.0, y)
        return y
    return mv


@pytest.mark.parametrize("name", ["d_N", "d_SL", "z_N_helmholtz", "z_HL"])
def test_fixed_iterations_match_the_oracle(capi, name):
    from oracle.gmres_oracle import gmres

    flat, _, _ = load_golden(name)
    assert flat.nb_rows == flat.nb_cols
    op = capi.Operator(flat.desc)
    rng = np.random.default_rng(2)
    b = rnd(rng, flat.nb_rows, flat.np_dtype)
    mv = _shifted(flat, 0.0)
    for iters, restart, cgs2 in [(3, 40, True), (6, 2, True), (4, 40, False)]:  # short of full convergence: the iterates are comparable
        xo, io = gmres(mv, b, restart=restart, max_iterations=iters, tolerance=0.0, reorthogonalize=cgs2)
        xg = np.zeros_like(b)
        ig = op.gmres(b, xg, restart=restart, max_iterations=iters, tolerance=0.0, orthogonalization=capi.HTB_GMRES_CGS2 if cgs2 else capi.HTB_GMRES_CGS)
        assert ig["iterations"] == io["iterations"] == iters and ig["matvecs"] == io["matvecs"]
        assert abs(ig["relative_residual"] - io["relative_residual"]) <= 1e-9 * max(1.0, io["relative_residual"])
        assert abs(ig["true_relative_residual"] - io["true_relative_residual"]) <= 1e-9 * max(1.0, io["true_relative_residual"])
        assert rel_err(xg, xo) < 1e-8, (name, iters, restart, rel_err(xg, xo))
    op.close()


def test_live_reference_solve_converges(capi, have_ref):
    """Laplace 1/(1e-5 + 4 pi r) (diagonal 1e5: well conditioned), assembled by the reference: GMRES with HPDDM's defaults."""
    if not have_ref:
        pytest.skip("oracle/_ref did not travel with the repo")
    from oracle import refharness as R
    from oracle.gmres_oracle import gmres

    case = R.RefCase(n=20000, kernel="laplace_reg")
    op = capi.Operator(case.desc)
    rng = np.random.default_rng(4)
    b = rnd(rng, case.nb_rows, case.np_dtype)

    def mv(v):
        y = np.zeros_like(v)
        case.vector_product("N", 1.0, np.ascontiguousarray(v), 0.0, y, variant="openmp")
        return y

    xg = np.zeros_like(b)
    ig = op.gmres(b, xg)  # defaults: restart 40, 100 iterations, 1e-6, CGS
    assert ig["converged"] and ig["true_relative_residual"] < 2e-6
    assert np.linalg.norm(b - mv(xg)) / np.linalg.norm(b) < 2e-6  # residual with the REFERENCE's product
    xo, io = gmres(mv, b)
    assert io["converged"] and io["iterations"] == ig["iterations"]
    assert rel_err(xg, xo) < 1e-8
    # device pointers, initial guess = the solution: converged without iterating
    import torch

    b_d, x_d = torch.from_numpy(b).cuda(), torch.from_numpy(xg.copy()).cuda()
    ig2 = op.gmres(b_d.data_ptr(), x_d.data_ptr(), mem_kind=capi.HTB_MEM_DEVICE, tolerance=1e-5)
    assert ig2["converged"] and ig2["iterations"] == 0
    op.close()


def test_rejects_non_square(capi):
    flat, _, _ = load_golden("d_rect")
    op = capi.Operator(flat.desc)
    with pytest.raises(capi.HtbError):
        op.gmres(np.zeros(flat.nb_rows), np.zeros(flat.nb_rows))
    op.close()
