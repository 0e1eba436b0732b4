"""tools/make_golden.py — writes tests/golden/*.npz: small H-matrices assembled by the UNMODIFIED reference
(through oracle/_ref, see oracle/ref/ref_harness.hpp) together with the reference's own products on them.

Run in the build container (needs /root/reference to have built oracle/_ref):  python tools/make_golden.py
Each file holds the flattened leaves (oracle/flatcase.py format) and, per entry k, the inputs and the
reference outputs:  trans, mu, alpha, beta, x, y_in, y_seq (sequential_internal_*), y_omp (openmp_internal_*).
Random inputs come from numpy default_rng(seed) — the reference's own tests use std::random_device
(include/htool/testing/generator_input.hpp:18,67) and are not reproducible, so the vectors are frozen here.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import refharness as R  # noqa: E402
from oracle.flatcase import FlatCase  # noqa: E402

CASES = {
    "d_N": dict(n=300),
    "d_SL": dict(n=300, symmetry="S", uplo="L"),
    "d_SU": dict(n=280, symmetry="S", uplo="U"),
    "z_HU": dict(n=260, dtype="complex", kernel="hermitian_reg", symmetry="H", uplo="U"),
    "z_HL": dict(n=240, dtype="complex", kernel="hermitian_reg", symmetry="H", uplo="L"),
    "z_SL": dict(n=250, dtype="complex", kernel="complex_reg", symmetry="S", uplo="L"),
    "z_N_helmholtz": dict(n=240, dtype="complex", kernel="helmholtz", leaf_size=40, geometry="disk"),
    "d_rect": dict(n=300, n_source=220, same_cluster=False, geometry="disk", z_source=0.3, kernel="laplace", epsilon=1e-6),
    "d_strip_SU": dict(n=540, n_partitions=3, partition_rank=1, symmetry="S", uplo="U"),
    "d_strip_N": dict(n=480, n_partitions=3, partition_rank=2),
    "z_strip_SL": dict(n=360, n_partitions=2, partition_rank=1, dtype="complex", kernel="helmholtz", symmetry="S", uplo="L", leaf_size=30, geometry="disk"),
    "d_SVD": dict(n=280, compressor="SVD", epsilon=1e-4, leaf_size=30),
    "d_bigleaf_SL": dict(n=260, leaf_size=90, epsilon=1e-8, symmetry="S", uplo="L"),
}


def rnd(rng, n, dt):
    v = rng.random(n) - 0.5
    if dt == np.complex128:
        v = v + 1j * (rng.random(n) - 0.5)
    return v.astype(dt)


def main():
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, kw in CASES.items():
        case = R.RefCase(**kw)
        flat = FlatCase.from_desc(case.desc)
        info = case.info()
        sym = info["symmetry_for_leaves"]
        rng = np.random.default_rng(abs(hash(name)) % (2**31) if False else sum(map(ord, name)))
        arrays = flat.save_arrays()
        k = 0
        for trans in "NTC":
            if (trans == "T" and sym == "H") or (trans == "C" and sym == "S"):
                continue
            ni, no = (case.nb_cols, case.nb_rows) if trans == "N" else (case.nb_rows, case.nb_cols)
            for mu, (alpha, beta) in [(1, (1.0, 0.0)), (1, (0.75, -1.25)), (3, (1.5, 0.5))]:
                if case.np_dtype == np.complex128:
                    alpha, beta = alpha * (1 + 0.5j), beta * (1 - 0.25j)
                x, y_in = rnd(rng, ni * mu, case.np_dtype), rnd(rng, no * mu, case.np_dtype)
                y_seq, y_omp = y_in.copy(), y_in.copy()
                if mu == 1:
                    case.vector_product(trans, alpha, x, beta, y_seq, variant="sequential")
                    case.vector_product(trans, alpha, x, beta, y_omp, variant="openmp")
                else:
                    case.matrix_product_row_major(trans, alpha, x, beta, y_seq, mu, variant="sequential")
                    case.matrix_product_row_major(trans, alpha, x, beta, y_omp, mu, variant="openmp")
                arrays[f"e{k}_meta"] = np.array([ord(trans), mu], dtype=np.int64)
                arrays[f"e{k}_ab"] = np.array([alpha, beta], dtype=case.np_dtype)
                arrays[f"e{k}_x"], arrays[f"e{k}_yin"], arrays[f"e{k}_yseq"], arrays[f"e{k}_yomp"] = x, y_in, y_seq, y_omp
                k += 1
        arrays["n_entries"] = np.array([k], dtype=np.int64)
        arrays["perm_target"], arrays["perm_source"] = case.permutation(0), case.permutation(1)
        # user-numbering product (add_hmatrix_vector_product, exec par) for the permutation front end
        # (only for whole operators: a row strip's permutation is not local, cluster_node.hpp:150-175 rejects it)
        if kw.get("partition_rank", -1) < 0:
            xu, yu = rnd(rng, case.nb_cols, case.np_dtype), rnd(rng, case.nb_rows, case.np_dtype)
            yu_ref = yu.copy()
            ab = (0.5, 2.0)
            case.vector_product("N", ab[0], xu, ab[1], yu_ref, variant="user")
            arrays["user_x"], arrays["user_yin"], arrays["user_yref"] = xu, yu, yu_ref
            arrays["user_ab"] = np.array(ab, dtype=case.np_dtype)
            mu = 2
            Xu, Yu = rnd(rng, case.nb_cols * mu, case.np_dtype), rnd(rng, case.nb_rows * mu, case.np_dtype)
            Yu_ref = Yu.copy()
            case.matrix_product_user("N", ab[0], Xu, ab[1], Yu_ref, mu)  # column-major B (n x mu), C (m x mu)
            arrays["userm_x"], arrays["userm_yin"], arrays["userm_yref"] = Xu, Yu, Yu_ref
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {info['nb_rows']}x{info['nb_cols']} leaves={info['nb_leaves']} (dense {info['nb_dense_leaves']}, lr {info['nb_low_rank_leaves']}, twice {info['nb_leaves_applied_twice']})"
              f" coeffs={info['coefficients']} rank {info['rank_min']}..{info['rank_max']} entries={k} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
