"""tools/make_golden_aca.py — writes tests/golden/aca/*.npz: small H-matrices ASSEMBLED by the UNMODIFIED reference
(HMatrixTreeBuilder with its default compressor sympartialACA, through oracle/_ref) together with what the device assembly
needs to redo the work: the kernel function, epsilon and the points of the root block in cluster numbering.

Run in the build container (needs /root/reference to have built oracle/_ref):  python tools/make_golden_aca.py
Each file holds the flattened leaves (oracle/flatcase.py format: the reference's ranks and U / V factors, and the dense
leaves), `points_target`, `points_source` (n x 3), `aca_meta` = [kernel id (capi.HTB_KERNELS), epsilon, wavenumber] and one reference
product (x, y = H x by openmp_internal_add_hmatrix_vector_product).
The blocks of a fixture were compressed with the BLAS of this image (OpenBLAS: daxpy = rounded product + rounded sum); the
oracle (oracle/aca_oracle.c, fma_axpy = 0) and the device kernels (option aca_fma_axpy = 0) reproduce them bit for bit.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from htool_b200.capi import HTB_KERNELS  # noqa: E402
from oracle import refharness as R  # noqa: E402
from oracle.flatcase import FlatCase  # noqa: E402

CASES = {
    "d_N": dict(n=520, kernel="laplace_reg", epsilon=1e-4),
    "d_SL": dict(n=640, kernel="laplace_reg", epsilon=1e-4, symmetry="S", uplo="L"),
    "d_SU_eps6": dict(n=520, kernel="laplace_reg", epsilon=1e-6, symmetry="S", uplo="U"),
    "d_rect": dict(n=600, n_source=450, same_cluster=False, z_source=1.5, kernel="laplace", epsilon=1e-4),
    "d_strip": dict(n=900, n_partitions=3, partition_rank=1, kernel="laplace_reg", epsilon=1e-3),
    # complex kernel functions (the reference's template instantiated for std::complex<double>)
    "z_SL": dict(n=420, dtype="complex", kernel="complex_reg", epsilon=1e-4, symmetry="S", uplo="L"),
    "z_HU": dict(n=400, dtype="complex", kernel="hermitian_reg", epsilon=1e-4, symmetry="H", uplo="U"),
    "z_helmholtz": dict(n=360, dtype="complex", kernel="helmholtz", epsilon=1e-4, wavenumber=5.0),
}


def main():
    out_dir = os.path.join(REPO, "tests", "golden", "aca")
    os.makedirs(out_dir, exist_ok=True)
    for name, kw in CASES.items():
        case = R.RefCase(**kw)
        flat = FlatCase.from_desc(case.desc)
        info = case.info()
        arrays = flat.save_arrays()
        arrays["points_target"], arrays["points_source"] = case.points(0), case.points(1)
        arrays["aca_meta"] = np.array([HTB_KERNELS[kw["kernel"]], kw["epsilon"], kw.get("wavenumber", 0.0)], dtype=np.float64)
        rng = np.random.default_rng(sum(map(ord, name)))
        x = (rng.random(case.nb_cols) - 0.5).astype(case.np_dtype)
        if case.np_dtype == np.complex128:
            x = x + 1j * (rng.random(case.nb_cols) - 0.5)
        y = np.zeros(case.nb_rows, case.np_dtype)
        case.vector_product("N", 1.0, x, 0.0, y, variant="openmp")
        arrays["x"], arrays["y"] = x, y
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {info['nb_rows']}x{info['nb_cols']} leaves={info['nb_leaves']} (dense {info['nb_dense_leaves']}, lr {info['nb_low_rank_leaves']})"
              f" coeffs={info['coefficients']} rank {info['rank_min']}..{info['rank_max']} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
