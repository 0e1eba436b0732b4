"""oracle/refharness.py — TEST INFRASTRUCTURE. ctypes wrapper of oracle/_ref/libhtool_ref.so, i.e. the
UNMODIFIED reference (htool headers under /root/reference/include, compiled by oracle/Makefile) behind the
small C API of oracle/ref/ref_capi.cpp.

Used by tests/ (parity checker), tools/make_golden.py (fixture generation) and bench.py (the reference
assembles the workload on the host, as the north_star prescribes, and its OpenMP product is the
cpu_baseline / `--impl reference` arm). The product never imports this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from htool_b200.capi import LEAF_NP_DTYPE, htb_hmatrix_desc, htb_leaf

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libhtool_ref.so")

KERNELS = {"laplace": 0, "laplace_reg": 1, "complex_reg": 2, "hermitian_reg": 3, "helmholtz": 4, "complex": 5}
GEOMETRIES = {"sphere_surface": 0, "ball": 1, "disk": 2}
COMPRESSORS = {"sympartialACA": 0, "SVD": 1, "fullACA": 2, "partialACA": 3}
VARIANTS = {"openmp": 0, "sequential": 1, "user": 2, "local_to_local_operator": 3, "global_to_local_operator": 4}


class ref_case_spec(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("kernel", C.c_int32),
        ("geometry_target", C.c_int32),
        ("geometry_source", C.c_int32),
        ("n_target", C.c_int32),
        ("n_source", C.c_int32),
        ("same_cluster", C.c_int32),
        ("min_depth", C.c_int32),
        ("leaf_size", C.c_int32),
        ("n_partitions", C.c_int32),
        ("partition_rank", C.c_int32),
        ("local_block", C.c_int32),
        ("compressor", C.c_int32),
        ("symmetry", C.c_int32),
        ("uplo", C.c_int32),
        ("reserved", C.c_int32),
        ("z_target", C.c_double),
        ("z_source", C.c_double),
        ("epsilon", C.c_double),
        ("eta", C.c_double),
        ("wavenumber", C.c_double),
    ]


_lib = None


def available() -> bool:
    return os.path.exists(REF_LIB)


def load():
    global _lib
    if _lib is None:
        # parallelism is over leaves (OpenMP); the bundled OpenBLAS must stay single-threaded
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
        lib = C.CDLL(REF_LIB)
        try:
            C.CDLL(None).openblas_set_num_threads(1)
        except Exception:
            pass
        lib.ref_case_create.restype = C.c_void_p
        lib.ref_case_create.argtypes = [C.POINTER(ref_case_spec)]
        lib.ref_case_destroy.argtypes = [C.c_void_p]
        lib.ref_case_desc.restype = C.POINTER(htb_hmatrix_desc)
        lib.ref_case_desc.argtypes = [C.c_void_p]
        lib.ref_case_info.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_case_permutation.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_case_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_case_vector_product.argtypes = [C.c_void_p, C.c_int, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_case_matrix_product_row_major.argtypes = [C.c_void_p, C.c_int, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_case_matrix_product_user.argtypes = [C.c_void_p, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_case_dense_product.argtypes = [C.c_void_p, C.c_char, C.c_void_p, C.c_void_p]
        for f in ("ref_case_hmatrix", "ref_case_target_cluster", "ref_case_source_cluster"):
            getattr(lib, f).restype = C.c_void_p
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.ref_set_num_threads.argtypes = [C.c_int]
        lib.ref_get_max_threads.restype = C.c_int
        lib.ref_set_log_level.argtypes = [C.c_int]
        _lib = lib
    return _lib


INFO_KEYS = [
    "nb_rows", "nb_cols", "row_offset", "col_offset", "nb_leaves", "nb_dense_leaves", "nb_low_rank_leaves",
    "nb_leaves_applied_twice", "coefficients", "coefficients_twice", "rank_min", "rank_max", "symmetry_for_leaves",
    "uplo_for_leaves", "build_seconds", "cluster_seconds", "n_target", "n_source", "omp_threads",
]


def make_spec(*, dtype="double", kernel="laplace_reg", n=1000, n_source=None, geometry="sphere_surface",
              geometry_source=None, same_cluster=True, z_target=0.0, z_source=0.0, epsilon=1e-4, eta=10.0,
              symmetry="N", uplo="N", min_depth=0, leaf_size=0, n_partitions=1, partition_rank=-1, local_block=False,
              compressor="sympartialACA", wavenumber=5.0) -> ref_case_spec:
    s = ref_case_spec()
    s.dtype = 0 if dtype in ("double", np.float64) else 1
    s.kernel = KERNELS[kernel]
    s.geometry_target = GEOMETRIES[geometry]
    s.geometry_source = GEOMETRIES[geometry_source or geometry]
    s.n_target = n
    s.n_source = n_source or n
    s.same_cluster = 1 if same_cluster else 0
    s.min_depth, s.leaf_size = min_depth, leaf_size
    s.n_partitions, s.partition_rank, s.local_block = n_partitions, partition_rank, 1 if local_block else 0
    s.compressor = COMPRESSORS[compressor]
    s.symmetry, s.uplo = ord(symmetry), ord(uplo)
    s.z_target, s.z_source, s.epsilon, s.eta, s.wavenumber = z_target, z_source, epsilon, eta, wavenumber
    return s


class RefCase:
    """An H-matrix assembled by the reference + the reference's CPU products on it."""

    def __init__(self, **kw):
        self.lib = load()
        s = make_spec(**kw)
        self.spec = s
        self.np_dtype = np.float64 if s.dtype == 0 else np.complex128
        self.handle = self.lib.ref_case_create(C.byref(s))
        if not self.handle:
            raise RuntimeError("reference harness failed to build the case")
        self.desc = self.lib.ref_case_desc(self.handle).contents
        self.nb_rows, self.nb_cols = self.desc.nb_rows, self.desc.nb_cols

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ref_case_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        out = np.zeros(24)
        self.lib.ref_case_info(self.handle, out.ctypes.data, out.size)
        d = {k: (float(v) if k.endswith("seconds") else int(v)) for k, v in zip(INFO_KEYS, out)}
        d["symmetry_for_leaves"] = chr(d["symmetry_for_leaves"])
        d["uplo_for_leaves"] = chr(d["uplo_for_leaves"])
        return d

    def leaves(self) -> np.ndarray:
        """Structured numpy VIEW of the htb_leaf array owned by the case."""
        n = self.desc.nb_leaves
        addr = C.cast(self.desc.leaves, C.c_void_p).value
        buf = (C.c_char * (n * LEAF_NP_DTYPE.itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=LEAF_NP_DTYPE)

    def permutation(self, side: int) -> np.ndarray:
        out = np.zeros(self.nb_rows if side == 0 else self.nb_cols, dtype=np.int32)
        self.lib.ref_case_permutation(self.handle, side, out.ctypes.data)
        return out

    def points(self, side: int) -> np.ndarray:
        """(n, 3) points of the root block's rows (side 0) / columns (side 1) in cluster numbering."""
        out = np.zeros((self.nb_rows if side == 0 else self.nb_cols, 3))
        self.lib.ref_case_points(self.handle, side, out.ctypes.data)
        return out

    def desc_without_dense_data(self):
        """A copy of the descriptor whose dense leaves carry no coefficients (data0 = NULL): the input of
        htb_create_generated. Returns (desc, keepalive)."""
        lv = self.leaves().copy()
        lv["data0"][lv["rank"] < 0] = 0
        arr = (htb_leaf * max(1, len(lv))).from_buffer_copy(lv.tobytes() if len(lv) else bytes(C.sizeof(htb_leaf)))
        d = htb_hmatrix_desc()
        C.memmove(C.byref(d), C.byref(self.desc), C.sizeof(htb_hmatrix_desc))
        d.leaves = C.cast(arr, C.POINTER(htb_leaf))
        return d, arr

    def _sc(self, v):
        return np.array([v], dtype=self.np_dtype)

    def vector_product(self, trans, alpha, x, beta, y, variant="openmp"):
        a, b = self._sc(alpha), self._sc(beta)
        assert x.dtype == self.np_dtype and y.dtype == self.np_dtype
        self.lib.ref_case_vector_product(self.handle, VARIANTS[variant], trans.encode(), a.ctypes.data, x.ctypes.data, b.ctypes.data, y.ctypes.data)
        return y

    def matrix_product_row_major(self, trans, alpha, x, beta, y, mu, variant="openmp"):
        a, b = self._sc(alpha), self._sc(beta)
        assert x.dtype == self.np_dtype and y.dtype == self.np_dtype
        self.lib.ref_case_matrix_product_row_major(self.handle, VARIANTS[variant], trans.encode(), a.ctypes.data, x.ctypes.data, b.ctypes.data, y.ctypes.data, mu)
        return y

    def matrix_product_user(self, trans, alpha, x, beta, y, mu):
        a, b = self._sc(alpha), self._sc(beta)
        self.lib.ref_case_matrix_product_user(self.handle, trans.encode(), a.ctypes.data, x.ctypes.data, b.ctypes.data, y.ctypes.data, mu)
        return y

    def dense_product(self, trans, x):
        y = np.zeros(self.nb_rows if trans == "N" else self.nb_cols, dtype=self.np_dtype)
        self.lib.ref_case_dense_product(self.handle, trans.encode(), x.ctypes.data, y.ctypes.data)
        return y


def set_num_threads(n: int):
    load().ref_set_num_threads(n)


def max_threads() -> int:
    return load().ref_get_max_threads()
