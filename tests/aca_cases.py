"""Helpers shared by tests/test_aca_oracle.py (CPU) and tests/test_gpu_compress.py (GPU): the ACA fixtures of
tests/golden/aca/ (tools/make_golden_aca.py: H-matrices assembled by the unmodified reference) and the descriptors the
device assembly is given (leaves without coefficients)."""
import ctypes as C
import glob
import os

import numpy as np
from conftest import GOLDEN_DIR

from htool_b200 import capi
from htool_b200.capi import HTB_RANK_COMPRESS, LEAF_NP_DTYPE, htb_hmatrix_desc, htb_leaf
from oracle.flatcase import FlatCase, oracle_sympartial_aca

ACA_DIR = os.path.join(GOLDEN_DIR, "aca")
ACA_GOLDEN = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(ACA_DIR, "*.npz")))
KERNEL_NAMES = {v: k for k, v in capi.HTB_KERNELS.items()}
DIAG_FLAGS = capi.HTB_LEAF_DIAG_SYMMETRIC | capi.HTB_LEAF_DIAG_HERMITIAN


class AcaCase:
    def __init__(self, flat, tp, sp, kernel, epsilon, x=None, y=None, wavenumber=0.0):
        self.flat, self.tp, self.sp, self.kernel, self.epsilon, self.x, self.y = flat, np.ascontiguousarray(tp), np.ascontiguousarray(sp), kernel, float(epsilon), x, y
        self.wavenumber = float(wavenumber)
        self.dtype = flat.np_dtype

    @classmethod
    def golden(cls, name):
        z = np.load(os.path.join(ACA_DIR, name + ".npz"))
        meta = z["aca_meta"]
        return cls(FlatCase.from_arrays(z), z["points_target"], z["points_source"], KERNEL_NAMES[int(meta[0])], meta[1], z["x"], z["y"], meta[2] if len(meta) > 2 else 0.0)

    @classmethod
    def live(cls, **kw):
        """Assembled by the reference, live (oracle/_ref)."""
        from oracle import refharness as R

        case = R.RefCase(**kw)
        return cls(FlatCase.from_desc(case.desc), case.points(0), case.points(1), kw.get("kernel", "laplace_reg"), kw.get("epsilon", 1e-4), wavenumber=kw.get("wavenumber", 5.0 if kw.get("kernel") == "helmholtz" else 0.0))

    def factors(self, i):
        """U (m x r) and V (r x n) of low-rank leaf i as the reference stored them."""
        f = self.flat
        m, n, r = (int(v) for v in f.table[i, 2:5])
        s = int(f.starts[i])
        return f.coeffs[s: s + m * r].reshape(r, m).T, f.coeffs[s + m * r: s + m * r + r * n].reshape(n, r).T

    def oracle_block(self, i, fma_axpy=False):
        f = self.flat
        lr, lc, m, n = (int(v) for v in f.table[i, 0:4])
        return oracle_sympartial_aca(self.kernel, self.tp, self.sp, m, n, f.row_offset + lr, f.col_offset + lc, lr, lc, self.epsilon, fma_axpy, wavenumber=self.wavenumber)

    def stripped_desc(self, compress_mask=None):
        """The descriptor htb_create_compressed is given: no coefficients anywhere; leaves of `compress_mask` (default: the
        reference's low-rank leaves) are admissible blocks to compress, the others dense leaves to generate."""
        f = self.flat
        k = f.table.shape[0]
        if compress_mask is None:
            compress_mask = f.table[:, 4] >= 0
        arr = (htb_leaf * max(1, k))()
        view = np.frombuffer(arr, dtype=LEAF_NP_DTYPE)[:k]
        for j, name in enumerate(("row_offset", "col_offset", "nb_rows", "nb_cols", "rank", "flags")):
            view[name] = f.table[:, j]
        view["rank"] = np.where(compress_mask, HTB_RANK_COMPRESS, -1)
        view["data0"], view["data1"] = 0, 0
        d = htb_hmatrix_desc()
        C.memmove(C.byref(d), C.byref(f.desc), C.sizeof(htb_hmatrix_desc))
        d.leaves = C.cast(arr, C.POINTER(htb_leaf))
        return d, arr

    def reassembled(self, compress_mask, fma_axpy=False):
        """What the reference's builder would hold if the leaves of `compress_mask` were its admissible blocks: the ORACLE's
        sympartialACA on each of them (a dense leaf where it fails), the kernel function on the dense ones."""
        from test_generated_dense import kernel_values

        f = self.flat
        table = f.table.copy()
        parts = []
        for i in range(table.shape[0]):
            lr, lc, m, n = (int(v) for v in table[i, 0:4])
            q = -1
            if compress_mask[i]:
                q, U, V, _ = self.oracle_block(i, fma_axpy)
            if q > 0:
                table[i, 4] = q
                parts += [np.asfortranarray(U).ravel(order="F"), np.asfortranarray(V).ravel(order="F")]
            else:
                table[i, 4] = -1
                if f.table[i, 4] < 0:  # the reference's own dense data (symv leaves keep their unread triangle as stored)
                    s = int(f.starts[i])
                    parts.append(f.coeffs[s: s + m * n])
                else:
                    gi, gj = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
                    parts.append(kernel_values(self.kernel, self.tp[lr + gi], self.sp[lc + gj], self.wavenumber).astype(self.dtype).ravel(order="F"))
        coeffs = np.concatenate(parts) if parts else np.zeros(0, self.dtype)
        return FlatCase(f.dtype_code, f.nb_rows, f.nb_cols, f.row_offset, f.col_offset, f.symmetry, f.uplo, table, coeffs)


TASK_DT = np.dtype([("byte_off", "<u8"), ("lrow", "<i4"), ("lcol", "<i4"), ("p0", "<i4"), ("k0", "<i4"), ("h", "<u2"), ("w", "<u2"), ("ld", "<u2"), ("flags", "<u2")])


def packed_side(desc, side, generate):
    """Stream + dense tasks + low-rank tasks of one side (htb_pack_host)."""
    lib = capi.load()
    capi.set_option("pack_generate_dense", 1 if generate else 0)
    try:
        p = capi.htb_packed_side()
        capi.check(lib, lib.htb_pack_host(C.byref(desc), side, C.byref(p)))
        stream = np.frombuffer((C.c_char * p.stream_bytes).from_address(p.stream), dtype=np.uint8).copy() if p.stream_bytes else np.zeros(0, np.uint8)
        dense = np.frombuffer((C.c_char * (p.n_dense_tasks * 32)).from_address(p.dense_tasks), dtype=TASK_DT).copy() if p.n_dense_tasks else np.zeros(0, TASK_DT)
        lowrank = np.frombuffer((C.c_char * (p.n_lowrank_tasks * 32)).from_address(p.lowrank_tasks), dtype=TASK_DT).copy() if p.n_lowrank_tasks else np.zeros(0, TASK_DT)
        lib.htb_pack_free(C.byref(p))
    finally:
        capi.set_option("pack_generate_dense", 0)
    return stream, dense, lowrank
