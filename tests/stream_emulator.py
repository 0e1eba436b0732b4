"""tests/stream_emulator.py — a numpy interpreter of the packed leaf store, for the CPU test-suite.

It consumes the very bytes htb_create uploads (obtained through htb_pack_host, no GPU needed) with the same
unit / stage / block walk, pass sequence and index shifts as htool_b200/csrc/kernels.cu and capi.cu
(run_product). It checks the HOST logic — packer, stream format, scratch offsets, pass sequences — against
the oracle; it is not a product path (tests only, pure Python, slow) and it never runs in place of the CUDA
kernels: the `-m gpu` tests exercise those through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from htool_b200 import capi

BLOCK_DT = np.dtype([("row_start", "<i4"), ("nrows", "<i4"), ("first_stage", "<u4"), ("n_stages", "<u4"), ("flags", "<u4"), ("r0", "<u4"), ("r1", "<u4"), ("r2", "<u4")])
STAGE_DT = np.dtype([("byte_off", "<u8"), ("nbytes", "<u4"), ("flags", "<u4")])
COMBINE_DT = np.dtype([("dst", "<u4"), ("src", "<u4"), ("w", "<u4"), ("n_chunks", "<u4")])
UNIT_DT = np.dtype([("data_off", "<u4"), ("geom", "<u4"), ("aux_apply", "<u4"), ("aux_reduce", "<u4")])
UNIT_LOWRANK, UNIT_DENSE, UNIT_ADDVEC = 0, 1, 2


def _view(addr, count, dt):
    if count == 0:
        return np.zeros(0, dt)
    buf = (C.c_char * (count * dt.itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dt).copy()


class PackedSide:
    def __init__(self, desc, side):
        lib = capi.load()
        p = capi.htb_packed_side()
        capi.check(lib, lib.htb_pack_host(C.byref(desc), side, C.byref(p)))
        self.n, self.n_blocks = p.n, p.n_blocks
        self.scratch_elems = p.scratch_elems
        self.blocks = _view(p.blocks, p.n_blocks, BLOCK_DT)
        self.stages = _view(p.stages, p.n_stages, STAGE_DT)
        self.order = _view(p.order, p.n_blocks, np.dtype("<u4"))
        self.combine = _view(p.combine, p.n_combine, COMBINE_DT)
        self.stream = _view(p.stream, p.stream_bytes, np.dtype("u1"))
        self.stream_bytes = p.stream_bytes
        lib.htb_pack_free(C.byref(p))

    def units_of_stage(self, st, dtype):
        """Yields (unit record, panel as an (h, w) array) for one stage."""
        sd = self.stages[st]
        raw = self.stream[int(sd["byte_off"]): int(sd["byte_off"]) + int(sd["nbytes"])]
        n_units, data_off = np.frombuffer(raw[:8].tobytes(), dtype="<u4")
        units = np.frombuffer(raw[16:16 + 16 * int(n_units)].tobytes(), dtype=UNIT_DT)
        data = np.frombuffer(raw[int(data_off):].tobytes()[: (len(raw) - int(data_off)) // np.dtype(dtype).itemsize * np.dtype(dtype).itemsize], dtype=dtype)
        for u in units:
            g = int(u["geom"])
            row0, h, w, kind, twice = g & 0xFF, ((g >> 8) & 0xFF) + 1, (g >> 16) & 0xFF, (g >> 24) & 3, (g >> 26) & 1
            panel = None
            if kind != UNIT_ADDVEC:
                panel = data[int(u["data_off"]): int(u["data_off"]) + h * w].reshape(w, h).T  # column-major, ld = h
            yield u, row0, h, w, kind, twice, panel


class Emulator:
    """run_product of capi.cu, in numpy."""

    def __init__(self, flatcase):
        self.fc = flatcase
        self.dtype = flatcase.np_dtype
        self.side = [PackedSide(flatcase.desc, 0), PackedSide(flatcase.desc, 1)]
        self.scratch_elems = self.side[0].scratch_elems
        self.sym = flatcase.symmetry
        self.nb_rows, self.nb_cols = flatcase.nb_rows, flatcase.nb_cols
        self.D = flatcase.row_offset - flatcase.col_offset
        self.any_twice = any(bool((s.blocks["flags"] & 1).any()) for s in self.side)

    # REDUCE pass (reduce_kernel)
    def reduce(self, s, vec, in_shift, scratch, twice_only, conj):
        side = self.side[s]
        for b in side.order:
            bd = side.blocks[b]
            if bd["n_stages"] == 0 or (twice_only and not (bd["flags"] & 1)):
                continue
            xin = np.zeros(int(bd["nrows"]), self.dtype)
            for i in range(int(bd["nrows"])):
                g = int(bd["row_start"]) + i + in_shift
                if 0 <= g < len(vec):
                    xin[i] = vec[g]
            for st in range(int(bd["first_stage"]), int(bd["first_stage"] + bd["n_stages"])):
                if twice_only and not (side.stages[st]["flags"] & 1):
                    continue
                for u, row0, h, w, kind, twice, panel in side.units_of_stage(st, self.dtype):
                    if kind == UNIT_ADDVEC or (twice_only and not twice):
                        continue
                    P = np.conj(panel) if conj else panel
                    scratch[int(u["aux_reduce"]): int(u["aux_reduce"]) + w] = P.T @ xin[row0: row0 + h]

    def combine(self, s, scratch, twice_only):
        for ce in self.side[s].combine:
            if twice_only and not (int(ce["n_chunks"]) & 0x80000000):
                continue
            nc, w, src, dst = int(ce["n_chunks"]) & 0x7FFFFFFF, int(ce["w"]), int(ce["src"]), int(ce["dst"])
            scratch[dst: dst + w] = scratch[src: src + nc * w].reshape(nc, w).sum(axis=0)

    # APPLY pass (apply_kernel)
    def apply(self, s, vec_in, in_shift, out, out_shift, scratch, alpha, beta, twice_only, conj):
        side = self.side[s]
        for b in side.order:
            bd = side.blocks[b]
            if twice_only and not (bd["flags"] & 1):
                continue
            acc = np.zeros(int(bd["nrows"]), self.dtype)
            for st in range(int(bd["first_stage"]), int(bd["first_stage"] + bd["n_stages"])):
                if twice_only and not (side.stages[st]["flags"] & 1):
                    continue
                for u, row0, h, w, kind, twice, panel in side.units_of_stage(st, self.dtype):
                    if twice_only and not twice:
                        continue
                    a = int(u["aux_apply"])
                    if kind == UNIT_ADDVEC:
                        acc[row0: row0 + h] += scratch[a: a + h]
                        continue
                    if kind == UNIT_LOWRANK:
                        c = scratch[a: a + w]
                    else:
                        c = np.zeros(w, self.dtype)
                        for k in range(w):
                            g = a + k + in_shift
                            if 0 <= g < len(vec_in):
                                c[k] = vec_in[g]
                    P = np.conj(panel) if conj else panel
                    acc[row0: row0 + h] += P @ c
            for i in range(int(bd["nrows"])):
                g = int(bd["row_start"]) + i + out_shift
                if 0 <= g < len(out):
                    out[g] = alpha * acc[i] + (0 if beta == 0 else beta * out[g])

    def vector_product(self, trans, alpha, x, beta, y):
        sym = self.sym
        if (trans == "T" and sym == "H") or (trans == "C" and sym == "S"):
            return 2
        twice = sym != "N" and self.any_twice
        is_complex = self.dtype == np.complex128
        T1 = np.zeros(max(1, self.scratch_elems), self.dtype)
        T2 = np.zeros(max(1, self.scratch_elems), self.dtype)
        D = self.D
        if trans == "N":
            self.reduce(1, x, 0, T1, False, False)
            if twice:
                self.reduce(0, x, D, T2, True, sym == "H" and is_complex)
                self.combine(0, T2, True)
            self.combine(1, T1, False)
            self.apply(0, x, 0, y, 0, T1, alpha, beta, False, False)
            if twice:
                self.apply(1, x, 0, y, -D, T2, alpha, 1.0, True, sym == "H" and is_complex)
        else:
            conj = trans == "C" and is_complex
            self.reduce(0, x, 0, T1, False, conj)
            self.combine(0, T1, False)
            if twice:
                self.reduce(1, x, -D, T2, True, False)
                self.combine(1, T2, True)
            self.apply(1, x, 0, y, 0, T1, alpha, beta, False, conj)
            if twice:
                self.apply(0, x, -D, y, D, T2, alpha, 1.0, True, False)
        return 0
