"""CPU tests of the GMRES oracle (oracle/gmres_oracle.py): the numpy restatement is pinned against a direct solve and
against scipy's GMRES on dense systems before tests/test_gpu_gmres.py trusts it as the checker of htb_gmres."""
import numpy as np
import pytest

from oracle.gmres_oracle import gmres


def _system(n, dtype, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) / np.sqrt(n)
    if dtype == np.complex128:
        A = A + 1j * rng.standard_normal((n, n)) / np.sqrt(n)
    A = A + 3.0 * np.eye(n)
    b = rng.standard_normal(n).astype(dtype)
    if dtype == np.complex128:
        b = b + 1j * rng.standard_normal(n)
    return A.astype(dtype), b


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("restart", [5, 40])
def test_oracle_gmres_matches_direct_solve(dtype, restart):
    A, b = _system(120, dtype, 3)
    x, info = gmres(lambda v: A @ v, b, restart=restart, max_iterations=400, tolerance=1e-12)
    assert info["converged"] and info["true_relative_residual"] < 1e-11
    ref = np.linalg.solve(A, b)
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-10


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_oracle_gmres_follows_scipy_iteration_by_iteration(dtype):
    """Same Krylov space, same minimisation: after k iterations (no restart) the residual norms agree with scipy's."""
    sp = pytest.importorskip("scipy.sparse.linalg")
    A, b = _system(80, dtype, 5)
    for k in (3, 7, 15):
        x, info = gmres(lambda v: A @ v, b, restart=k, max_iterations=k, tolerance=0.0)
        xs, _ = sp.gmres(A, b, restart=k, maxiter=1, rtol=0.0, atol=0.0)
        r1, r2 = np.linalg.norm(b - A @ x), np.linalg.norm(b - A @ xs)
        assert abs(r1 - r2) <= 1e-9 * np.linalg.norm(b), (k, r1, r2)
        assert info["iterations"] == k and abs(info["relative_residual"] - r1 / np.linalg.norm(b)) < 1e-9


def test_oracle_gmres_cgs2_and_initial_guess():
    A, b = _system(100, np.float64, 9)
    x0 = np.linalg.solve(A, b) + 1e-3
    x, info = gmres(lambda v: A @ v, b, x0=x0, restart=10, max_iterations=100, tolerance=1e-13, reorthogonalize=True)
    assert info["converged"] and info["iterations"] < 40
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-12
    # zero right-hand side: converged at once, x untouched
    x, info = gmres(lambda v: A @ v, np.zeros(100), x0=np.zeros(100))
    assert info["converged"] and info["iterations"] == 0 and not x.any()
