"""oracle/flatcase.py — TEST INFRASTRUCTURE. A flattened H-matrix held in numpy arrays (so it can be saved
as a golden fixture and travel to the GPU box without the reference), plus the ctypes binding of the plain-C
oracle (oracle/libhtb_oracle.so, restating the reference's sequential product, see hmat_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from htool_b200.capi import (HTB_COMPLEX_DOUBLE, HTB_DOUBLE, HTB_LEAF_APPLY_TRANSPOSED_TOO, HTB_LEAF_DIAG_HERMITIAN,
                             HTB_LEAF_DIAG_SYMMETRIC, HTB_LEAF_UPLO_UPPER, LEAF_NP_DTYPE, htb_hmatrix_desc, htb_leaf)

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "libhtb_oracle.so")

_lib = None


def load_oracle():
    global _lib
    if _lib is None:
        lib = C.CDLL(ORACLE_LIB)
        lib.oracle_add_vector_product.restype = C.c_int
        lib.oracle_add_vector_product.argtypes = [C.POINTER(htb_hmatrix_desc), C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_add_matrix_product_row_major.restype = C.c_int
        lib.oracle_add_matrix_product_row_major.argtypes = [C.POINTER(htb_hmatrix_desc), C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_sympartial_aca.restype = C.c_int
        lib.oracle_sympartial_aca.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_sympartial_aca_z.restype = C.c_int
        lib.oracle_sympartial_aca_z.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def oracle_sympartial_aca(kernel, target_points, source_points, m, n, row_offset, col_offset, lrow, lcol, epsilon, fma_axpy=False, max_rank=None, wavenumber=0.0):
    """oracle/aca_oracle.c on one block. kernel: 'laplace' | 'laplace_reg'. Returns (rank, U (m x rank, Fortran order),
    V (rank x n, Fortran order), pivots (rank x 2)); rank = -1: the reference reports a failure (dense leaf)."""
    lib = load_oracle()
    cap = max_rank or max(1, (m * n) // (m + n))
    tp, sp = np.ascontiguousarray(target_points, dtype=np.float64), np.ascontiguousarray(source_points, dtype=np.float64)
    piv = np.zeros(2 * cap, dtype=np.int32)
    if kernel in ("laplace", "laplace_reg"):
        U, V = np.zeros(m * cap), np.zeros(cap * n)
        q = lib.oracle_sympartial_aca({"laplace": 0, "laplace_reg": 1}[kernel], tp.ctypes.data, sp.ctypes.data, m, n, row_offset, col_offset, lrow, lcol, float(epsilon), int(bool(fma_axpy)), cap,
                                      U.ctypes.data, V.ctypes.data, piv.ctypes.data)
    else:  # complex kernel functions (ids of oracle/ref/ref_harness.hpp)
        U, V = np.zeros(m * cap, np.complex128), np.zeros(cap * n, np.complex128)
        q = lib.oracle_sympartial_aca_z({"complex_reg": 2, "hermitian_reg": 3, "helmholtz": 4, "complex": 5}[kernel], float(wavenumber), tp.ctypes.data, sp.ctypes.data, m, n, row_offset, col_offset, lrow, lcol,
                                        float(epsilon), int(bool(fma_axpy)), cap, U.ctypes.data, V.ctypes.data, piv.ctypes.data)
    if q == -2:
        raise RuntimeError("max_rank too small")
    if q <= 0:
        return q, None, None, None
    return q, U[: m * q].reshape(q, m).T, V[: q * n].reshape(n, q).T, piv[: 2 * q].reshape(q, 2)


class FlatCase:
    """Leaf table + one contiguous coefficient buffer. `table` columns: row_offset, col_offset, nb_rows,
    nb_cols, rank, flags (int32). Coefficients of leaf i start at element `starts[i]`: dense A (m*n,
    column-major) or U (m*r column-major) followed by V (r*n column-major)."""

    def __init__(self, dtype_code, nb_rows, nb_cols, row_offset, col_offset, symmetry, uplo, table, coeffs):
        self.dtype_code = int(dtype_code)
        self.np_dtype = np.float64 if self.dtype_code == HTB_DOUBLE else np.complex128
        self.nb_rows, self.nb_cols = int(nb_rows), int(nb_cols)
        self.row_offset, self.col_offset = int(row_offset), int(col_offset)
        self.symmetry, self.uplo = symmetry, uplo
        self.table = np.ascontiguousarray(table, dtype=np.int32).reshape(-1, 6)
        self.coeffs = np.ascontiguousarray(coeffs, dtype=self.np_dtype)
        m, n, r = self.table[:, 2].astype(np.int64), self.table[:, 3].astype(np.int64), self.table[:, 4].astype(np.int64)
        self.sizes0 = np.where(r < 0, m * n, m * np.maximum(r, 0))
        self.sizes1 = np.where(r < 0, 0, np.maximum(r, 0) * n)
        tot = self.sizes0 + self.sizes1
        self.starts = np.concatenate([[0], np.cumsum(tot)[:-1]]) if len(tot) else np.zeros(0, np.int64)
        assert int(tot.sum()) == self.coeffs.size, (int(tot.sum()), self.coeffs.size)
        self._build_desc()

    # -- construction ----------------------------------------------------------------------------------
    @classmethod
    def from_desc(cls, desc: htb_hmatrix_desc):
        """Deep copy out of a live descriptor (e.g. RefCase.desc)."""
        n = desc.nb_leaves
        dtype = np.float64 if desc.dtype == HTB_DOUBLE else np.complex128
        addr = C.cast(desc.leaves, C.c_void_p).value
        lv = np.frombuffer((C.c_char * (n * LEAF_NP_DTYPE.itemsize)).from_address(addr), dtype=LEAF_NP_DTYPE) if n else np.zeros(0, LEAF_NP_DTYPE)
        table = np.stack([lv[k] for k in ("row_offset", "col_offset", "nb_rows", "nb_cols", "rank", "flags")], axis=1) if n else np.zeros((0, 6), np.int32)
        parts = []
        isz = np.dtype(dtype).itemsize
        for i in range(n):
            m, nn, r = int(lv["nb_rows"][i]), int(lv["nb_cols"][i]), int(lv["rank"][i])
            s0 = m * nn if r < 0 else m * max(r, 0)
            s1 = 0 if r < 0 else max(r, 0) * nn
            if s0:
                parts.append(np.frombuffer((C.c_char * (s0 * isz)).from_address(int(lv["data0"][i])), dtype=dtype))
            if s1:
                parts.append(np.frombuffer((C.c_char * (s1 * isz)).from_address(int(lv["data1"][i])), dtype=dtype))
        coeffs = np.concatenate(parts) if parts else np.zeros(0, dtype)
        return cls(desc.dtype, desc.nb_rows, desc.nb_cols, desc.row_offset, desc.col_offset,
                   desc.symmetry_for_leaves.decode(), desc.uplo_for_leaves.decode(), table, coeffs)

    def save_arrays(self, prefix=""):
        return {
            prefix + "meta": np.array([self.dtype_code, self.nb_rows, self.nb_cols, self.row_offset, self.col_offset, ord(self.symmetry), ord(self.uplo)], dtype=np.int64),
            prefix + "table": self.table,
            prefix + "coeffs": self.coeffs,
        }

    @classmethod
    def from_arrays(cls, z, prefix=""):
        meta = z[prefix + "meta"]
        return cls(meta[0], meta[1], meta[2], meta[3], meta[4], chr(meta[5]), chr(meta[6]), z[prefix + "table"], z[prefix + "coeffs"])

    def _build_desc(self):
        k = self.table.shape[0]
        self._leaves = (htb_leaf * max(k, 1))()
        if k:
            view = np.frombuffer(self._leaves, dtype=LEAF_NP_DTYPE)[:k]
            for j, name in enumerate(("row_offset", "col_offset", "nb_rows", "nb_cols", "rank", "flags")):
                view[name] = self.table[:, j]
            base = self.coeffs.ctypes.data
            isz = self.coeffs.itemsize
            view["data0"] = (base + self.starts * isz).astype(np.uint64)
            d1 = (base + (self.starts + self.sizes0) * isz).astype(np.uint64)
            view["data1"] = np.where(self.table[:, 4] < 0, 0, d1)
        d = htb_hmatrix_desc()
        d.dtype = self.dtype_code
        d.nb_rows, d.nb_cols = self.nb_rows, self.nb_cols
        d.row_offset, d.col_offset = self.row_offset, self.col_offset
        d.symmetry_for_leaves = self.symmetry.encode()
        d.uplo_for_leaves = self.uplo.encode()
        d.device = -1
        d.nb_leaves = k
        d.leaves = C.cast(self._leaves, C.POINTER(htb_leaf))
        self.desc = d

    # -- numbers ---------------------------------------------------------------------------------------
    @property
    def coefficients(self) -> int:
        return int(self.coeffs.size)

    @property
    def coefficients_twice(self) -> int:
        tw = (self.table[:, 5] & HTB_LEAF_APPLY_TRANSPOSED_TOO) != 0
        return int((self.sizes0 + self.sizes1)[tw].sum())

    # -- the plain-C oracle ----------------------------------------------------------------------------
    def _sc(self, v):
        return np.array([v], dtype=self.np_dtype)

    def oracle_vector_product(self, trans, alpha, x, beta, y):
        a, b = self._sc(alpha), self._sc(beta)
        assert x.dtype == self.np_dtype and y.dtype == self.np_dtype
        return load_oracle().oracle_add_vector_product(C.byref(self.desc), trans.encode(), a.ctypes.data, x.ctypes.data, b.ctypes.data, y.ctypes.data)

    def oracle_matrix_product_row_major(self, trans, alpha, x, beta, y, mu):
        a, b = self._sc(alpha), self._sc(beta)
        assert x.dtype == self.np_dtype and y.dtype == self.np_dtype
        return load_oracle().oracle_add_matrix_product_row_major(C.byref(self.desc), trans.encode(), a.ctypes.data, x.ctypes.data, b.ctypes.data, y.ctypes.data, mu)


def random_flatcase(seed=0, dtype_code=HTB_DOUBLE, nb_rows=300, nb_cols=260, n_leaves=60, max_dim=70, max_rank=9, symmetric=None):
    """A synthetic leaf list that needs no reference: random (overlapping) blocks with random payloads.
    It exercises the packer/kernels on shapes a block tree never produces (ragged, overlapping, rank 0,
    1 x 1, full-width). With symmetric in ('S','H') the operator is square, some off-diagonal leaves are
    applied twice and some diagonal dense leaves carry the symv/hemv flags."""
    rng = np.random.default_rng(seed)
    dtype = np.float64 if dtype_code == HTB_DOUBLE else np.complex128
    if symmetric:
        nb_cols = nb_rows
    rows, parts = [], []

    def rnd(k):
        v = rng.standard_normal(k)
        if dtype_code == HTB_COMPLEX_DOUBLE:
            v = v + 1j * rng.standard_normal(k)
        return v.astype(dtype)

    for i in range(n_leaves):
        m = int(rng.integers(1, min(max_dim, nb_rows) + 1))
        n = int(rng.integers(1, min(max_dim, nb_cols) + 1))
        if i == 0:
            m, n = nb_rows, nb_cols  # one leaf spanning everything
        r0 = int(rng.integers(0, nb_rows - m + 1))
        c0 = int(rng.integers(0, nb_cols - n + 1))
        flags = 0
        kind = rng.integers(0, 3)
        if symmetric and i % 5 == 1:
            n = m
            c0 = r0
            rank = -1
            flags = (HTB_LEAF_DIAG_SYMMETRIC if symmetric == "S" else HTB_LEAF_DIAG_HERMITIAN) | (HTB_LEAF_UPLO_UPPER if i % 2 else 0)
        elif kind == 0:
            rank = -1
        else:
            rank = int(rng.integers(0, max_rank + 1))
        if symmetric and not flags and c0 != r0 and i % 3 != 0:
            flags |= HTB_LEAF_APPLY_TRANSPOSED_TOO
        rows.append([r0, c0, m, n, rank, flags])
        if rank < 0:
            parts.append(rnd(m * n))
        else:
            parts.append(rnd(m * rank))
            parts.append(rnd(rank * n))
    coeffs = np.concatenate(parts) if parts else np.zeros(0, dtype)
    return FlatCase(dtype_code, nb_rows, nb_cols, 0, 0, symmetric or "N", "L" if symmetric else "N", np.array(rows, dtype=np.int32), coeffs)
