"""oracle/gmres_oracle.py — TEST INFRASTRUCTURE: numpy restatement of restarted GMRES, the checker of htb_gmres.

Reference side: DDM::solve with "-hpddm_schwarz_method none" (include/htool/solvers/ddm.hpp:134-193) hands the
DistributedOperator to HPDDM::IterativeMethod::solve, whose default method is GMRES (restart 40, tolerance 1e-6 relative
to ||b||, at most 100 iterations, classical Gram-Schmidt) with HPDDMOperator::GMV as the operator
(include/htool/wrappers/wrapper_hpddm.hpp:102-145). HPDDM is a third-party dependency that is NOT vendored in
/root/reference (cmake_modules/FindHPDDM.cmake:11-16; CI pins hpddm/hpddm@24aed69dbde7ef1526ae87ccf4f39ceb840bccea,
.github/workflows/CI.yml:134-137), so this file restates the published algorithm (Saad & Schultz, SIAM J. Sci. Stat.
Comput. 7, 1986: Arnoldi with Gram-Schmidt, Givens rotations on the Hessenberg matrix, restart) — PARITY UNPINNED with
respect to HPDDM's own arithmetic. It is pinned against numpy.linalg.solve and scipy.sparse.linalg.gmres on dense
systems (tests/test_gmres_oracle.py), and the anchor on the reference is the operator inside it: the matvec callable
the tests pass is the reference's / the oracle's own product.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def gmres(matvec, b, x0=None, restart=40, max_iterations=100, tolerance=1e-6, reorthogonalize=False):
    """Solves A x = b. matvec(v) -> A v. Returns (x, info) with info = dict(iterations, converged, matvecs,
    relative_residual [Givens recurrence], true_relative_residual)."""
    b = np.asarray(b)
    dtype = np.result_type(b.dtype, np.float64)
    n = b.size
    x = np.zeros(n, dtype) if x0 is None else np.array(x0, dtype=dtype, copy=True)
    bnorm = float(np.linalg.norm(b))
    denom = bnorm if bnorm > 0 else 1.0
    m = max(1, min(restart, max_iterations))
    it, converged, matvecs, rel = 0, False, 0, 1.0
    while it < max_iterations and not converged:
        r = b - matvec(x)
        matvecs += 1
        beta = float(np.linalg.norm(r))
        rel = beta / denom
        if rel <= tolerance or beta == 0.0:
            converged = True
            break
        V = np.zeros((m + 1, n), dtype)
        H = np.zeros((m + 1, m), dtype)
        cs, sn = np.zeros(m), np.zeros(m, dtype)
        g = np.zeros(m + 1, dtype)
        g[0] = beta
        V[0] = r / beta
        j = 0
        while j < m and it < max_iterations:
            it += 1
            w = matvec(V[j])
            matvecs += 1
            h = np.conj(V[: j + 1]) @ w  # classical Gram-Schmidt: all projections against the SAME w
            w = w - V[: j + 1].T @ h
            if reorthogonalize:
                h2 = np.conj(V[: j + 1]) @ w
                w = w - V[: j + 1].T @ h2
                h = h + h2
            H[: j + 1, j] = h
            hn = float(np.linalg.norm(w))
            H[j + 1, j] = hn
            if hn > 0:
                V[j + 1] = w / hn
            for k in range(j):
                t = cs[k] * H[k, j] + sn[k] * H[k + 1, j]
                H[k + 1, j] = -np.conj(sn[k]) * H[k, j] + cs[k] * H[k + 1, j]
                H[k, j] = t
            a, bb = abs(H[j, j]), abs(H[j + 1, j])
            t = float(np.hypot(a, bb))
            if t == 0.0:
                cs[j], sn[j] = 1.0, 0.0
            elif a == 0.0:
                cs[j], sn[j] = 0.0, np.conj(H[j + 1, j]) / bb
            else:
                cs[j] = a / t
                sn[j] = (H[j, j] / a) * np.conj(H[j + 1, j]) / t
            H[j, j] = cs[j] * H[j, j] + sn[j] * H[j + 1, j]
            H[j + 1, j] = 0.0
            g[j + 1] = -np.conj(sn[j]) * g[j]
            g[j] = cs[j] * g[j]
            rel = abs(g[j + 1]) / denom
            j += 1
            if rel <= tolerance or hn == 0.0:
                converged = rel <= tolerance
                break
        y = np.zeros(j, dtype)
        for k in range(j - 1, -1, -1):
            y[k] = (g[k] - H[k, k + 1: j] @ y[k + 1: j]) / H[k, k]
        x = x + V[:j].T @ y
    true_rel = float(np.linalg.norm(b - matvec(x))) / denom
    return x, dict(iterations=it, converged=bool(converged), matvecs=matvecs + 1, relative_residual=float(rel), true_relative_residual=true_rel)
