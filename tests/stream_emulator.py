"""tests/stream_emulator.py — a numpy interpreter of the packed leaf store, for the CPU test-suite.

It consumes the very bytes and tables htb_create uploads (obtained through htb_pack_host, no GPU needed) with
the same unit / stage / block walk, c-stream addressing, combine tables, pass sequence and index shifts as
htool_b200/csrc/kernels.cu and capi.cu (run_product). It checks the HOST logic — packer, stream format,
scratch offsets, pass sequences — against the oracle; it is not a product path (tests only, pure Python, slow)
and it never runs in place of the CUDA kernels: the `-m gpu` tests exercise those through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from htool_b200 import capi

BLOCK_DT = np.dtype([("row_start", "<i4"), ("nrows", "<i4"), ("first_stage", "<u4"), ("n_stages", "<u4"), ("flags", "<u4"), ("r0", "<u4"), ("r1", "<u4"), ("r2", "<u4")])
STAGE_DT = np.dtype([("byte_off", "<u8"), ("nbytes", "<u4"), ("c_off", "<u4"), ("c_len", "<u2"), ("flags", "<u2"), ("first_unit", "<u4"), ("n_units", "<u2"), ("n_panel", "<u2"), ("aux_off16", "<u4")])
COMBINE_DT = np.dtype([("src", "<u4"), ("dst_first", "<u4"), ("n_dst", "<u4"), ("packed", "<u4")])
COMBINE_DST_DT = np.dtype([("slot", "<u4"), ("sub_off", "<u2"), ("sub_len", "<u2")])
UNIT_DT = np.dtype([("data_off", "<u4"), ("geom", "<u4"), ("out", "<u4"), ("cslot", "<u2"), ("reserved", "<u2")])
MUNIT_DT = np.dtype([("out", "<u4"), ("src", "<u4"), ("poff", "<u4"), ("flags", "<u4")])
RUN_DT = np.dtype([("data_off", "<u4"), ("col0", "<u2"), ("K", "<u2"), ("row0", "u1"), ("h_minus_1", "u1"), ("flags", "u1"), ("r8", "u1"), ("r32", "<u4")])
assert RUN_DT.itemsize == 16
UNIT_LOWRANK, UNIT_DENSE, UNIT_ADDVEC = 0, 1, 2
PANEL_BUFFER_ELEMS = 4608
assert BLOCK_DT.itemsize == 32 and STAGE_DT.itemsize == 32 and COMBINE_DT.itemsize == 16 and COMBINE_DST_DT.itemsize == 8 and UNIT_DT.itemsize == 16


def unit_ld(h, isz):
    """store.hpp unit_ld: leading dimension of a panel of h rows (double: even)."""
    return (h + 1) & ~1 if isz == 8 else h


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
        self.cs_base, self.cs_elems, self.part_base, self.part_elems, self.piece_cols = p.cs_base, p.cs_elems, p.part_base, p.part_elems, p.piece_cols
        self.blocks = _view(p.blocks, p.n_blocks, BLOCK_DT)
        self.stages = _view(p.stages, p.n_stages, STAGE_DT)
        self.order = _view(p.order, p.n_blocks, np.dtype("<u4"))
        self.combine = _view(p.combine, p.n_combine, COMBINE_DT)
        self.combine_dst = _view(p.combine_dst, p.n_combine_dst, COMBINE_DST_DT)
        self.munits = _view(p.munits, p.n_munits, MUNIT_DT)
        self.combine_m = _view(p.combine_m, p.n_combine_m, COMBINE_DT)
        self.mscratch_elems, self.block_rows = p.mscratch_elems, p.block_rows
        self.stream = _view(p.stream, p.stream_bytes, np.dtype("u1"))
        self.stream_bytes = p.stream_bytes
        self.aux = {"reduce": _view(p.aux_reduce, p.aux_bytes, np.dtype("u1")), "apply": _view(p.aux_apply, p.aux_bytes, np.dtype("u1"))}
        lib.htb_pack_free(C.byref(p))

    def runs_of_stage(self, st, dtype, role):
        """The stage's aux record (store.hpp): yields (row0, h, twice, panel (h, K), column entries (K,)) per RUN, the
        panel being read from the stage bytes at the run's data offset, exactly as the multi-RHS kernels do."""
        sd = self.stages[st]
        raw = self.stream[int(sd["byte_off"]): int(sd["byte_off"]) + int(sd["nbytes"])]
        n_units, data_off, n_panel, _ = np.frombuffer(raw[:16].tobytes(), dtype="<u4")
        isz = np.dtype(dtype).itemsize
        data = np.frombuffer(raw[int(data_off):].tobytes()[: (len(raw) - int(data_off)) // isz * isz], dtype=dtype)
        aux_len = (int(sd["flags"]) >> 1) * 16
        rec = self.aux[role][int(sd["aux_off16"]) * 16: int(sd["aux_off16"]) * 16 + aux_len].tobytes()
        assert aux_len >= 16
        n_runs, n_cols = (int(v) for v in np.frombuffer(rec[:8], dtype="<u4"))
        assert aux_len == 16 + 16 * n_runs + 4 * n_cols and n_cols % 4 == 0
        runs = np.frombuffer(rec[16: 16 + 16 * n_runs], dtype=RUN_DT)
        cols = np.frombuffer(rec[16 + 16 * n_runs:], dtype="<u4")
        covered = 0
        for rd in runs:
            h, K = int(rd["h_minus_1"]) + 1, int(rd["K"])
            ld = unit_ld(h, isz)
            panel = data[int(rd["data_off"]): int(rd["data_off"]) + ld * K].reshape(K, ld).T[:h, :]
            covered += K
            yield int(rd["row0"]), h, int(rd["flags"]) & 1, panel, cols[int(rd["col0"]): int(rd["col0"]) + K]
        # the runs cover exactly the panel columns of the stage
        units = np.frombuffer(raw[16:16 + 16 * int(n_units)].tobytes(), dtype=UNIT_DT)[: int(n_panel)]
        assert covered == sum((int(u["geom"]) >> 16) & 0xFF for u in units)

    def units_of_stage(self, st, dtype):
        """Yields (unit record, row0, h, w, kind, twice, panel as an (h, w) array) for one stage."""
        sd = self.stages[st]
        raw = self.stream[int(sd["byte_off"]): int(sd["byte_off"]) + int(sd["nbytes"])]
        n_units, data_off, n_panel, first_unit = np.frombuffer(raw[:16].tobytes(), dtype="<u4")
        assert int(first_unit) == int(sd["first_unit"]) and int(n_units) == int(sd["n_units"]) and int(n_panel) == int(sd["n_panel"])
        units = np.frombuffer(raw[16:16 + 16 * int(n_units)].tobytes(), dtype=UNIT_DT)
        self.last_header = (int(n_units), int(n_panel), int(first_unit))
        isz = np.dtype(dtype).itemsize
        data = np.frombuffer(raw[int(data_off):].tobytes()[: (len(raw) - int(data_off)) // isz * isz], dtype=dtype)
        for u in units:
            g = int(u["geom"])
            row0, h, w, kind, twice = g & 0xFF, ((g >> 8) & 0xFF) + 1, (g >> 16) & 0xFF, (g >> 24) & 3, (g >> 26) & 1
            panel = None
            if kind != UNIT_ADDVEC:
                ld = unit_ld(h, isz)
                full = data[int(u["data_off"]): int(u["data_off"]) + ld * w].reshape(w, ld).T  # column-major
                assert not full[h:, :].any()  # pad rows are zero
                panel = full[:h, :]
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

    def new_scratch(self):
        # NaN-poisoned: a slot that is read must have been written by this product
        return np.full(max(2, self.scratch_elems), np.nan, self.dtype)

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
                    if twice_only and not twice:
                        continue
                    o = int(u["out"])
                    if kind == UNIT_ADDVEC:
                        scratch[o: o + h] = xin[row0: row0 + h]
                        continue
                    P = np.conj(panel) if conj else panel
                    scratch[o: o + w] = P.T @ xin[row0: row0 + h]

    # COMBINE pass of the direction whose consumer is side s (combine_kernel)
    def combine(self, s, scratch, twice_only):
        side = self.side[s]
        for ce in side.combine:
            pk = int(ce["packed"])
            n_sum, ln, tw = pk & 0xFFFFFF, (pk >> 24) & 0x7F, pk >> 31
            if twice_only and not tw:
                continue
            src = int(ce["src"])
            v = scratch[src: src + n_sum * ln].reshape(n_sum, ln).sum(axis=0)
            for d in side.combine_dst[int(ce["dst_first"]): int(ce["dst_first"]) + int(ce["n_dst"])]:
                so, sl = int(d["sub_off"]), int(d["sub_len"])
                scratch[int(d["slot"]): int(d["slot"]) + sl] = v[so: so + sl]

    # APPLY pass (apply_kernel)
    def apply(self, s, out, out_shift, scratch, alpha, beta, twice_only, conj):
        side = self.side[s]
        for b in side.order:
            bd = side.blocks[b]
            if twice_only and not (bd["flags"] & 1):
                continue
            acc = np.zeros(int(bd["nrows"]), self.dtype)
            for st in range(int(bd["first_stage"]), int(bd["first_stage"] + bd["n_stages"])):
                sd = side.stages[st]
                if twice_only and not (sd["flags"] & 1):
                    continue
                c0 = side.cs_base + int(sd["c_off"])
                assert c0 % (1 if self.dtype == np.complex128 else 2) == 0  # 16 B aligned bulk copy
                cseg = scratch[c0: c0 + int(sd["c_len"])]
                for u, row0, h, w, kind, twice, panel in side.units_of_stage(st, self.dtype):
                    if twice_only and not twice:
                        continue
                    a = int(u["cslot"])
                    if kind == UNIT_ADDVEC:
                        acc[row0: row0 + h] += cseg[a: a + h]
                        continue
                    P = np.conj(panel) if conj else panel
                    acc[row0: row0 + h] += P @ cseg[a: a + w]
            for i in range(int(bd["nrows"])):
                g = int(bd["row_start"]) + i + out_shift
                if 0 <= g < len(out):
                    out[g] = alpha * acc[i] + (0 if beta == 0 else beta * out[g])

    # ---- multi-RHS path (mkernels.cu / run_product_m of capi.cu): same streams, MUnit side tables --------------------
    def _munits_of_stage(self, s, st):
        """(unit tuple, MUnit) pairs of a stage, in stage order (panel units first)."""
        side = self.side[s]
        units = list(side.units_of_stage(st, self.dtype))
        n_units, n_panel, first = side.last_header
        kinds = [u[4] for u in units]
        assert all(k != UNIT_ADDVEC for k in kinds[:n_panel]) and all(k == UNIT_ADDVEC for k in kinds[n_panel:])
        return [(u, side.munits[first + i]) for i, u in enumerate(units)], n_panel

    def reduce_m(self, s, X, in_shift, M, twice_only, conj=False):
        """REDUCE_M: per RUN, T = op(panel)^T X[rows]; row k of the result goes to the scratch vector cols[k]."""
        side = self.side[s]
        for b in side.order:
            bd = side.blocks[b]
            if bd["n_stages"] == 0 or (twice_only and not (bd["flags"] & 1)):
                continue
            rows = int(bd["row_start"]) + np.arange(int(bd["nrows"])) + in_shift
            ok = (rows >= 0) & (rows < X.shape[0])
            xin = np.zeros((int(bd["nrows"]), X.shape[1]), self.dtype)
            xin[ok] = X[rows[ok]]
            for st in range(int(bd["first_stage"]), int(bd["first_stage"] + bd["n_stages"])):
                if twice_only and not (side.stages[st]["flags"] & 1):
                    continue
                # the unit view and the run view of a stage describe the same panels
                pairs, n_panel = self._munits_of_stage(s, st)
                unit_out = np.concatenate([int(mu["out"]) + np.arange(w) for (u, row0, h, w, kind, twice, panel), mu in pairs[:n_panel]] or [np.zeros(0, np.int64)])
                col = 0
                for row0, h, twice, panel, cols in side.runs_of_stage(st, self.dtype, "reduce"):
                    assert np.array_equal(cols, unit_out[col: col + len(cols)])
                    col += len(cols)
                    if twice_only and not twice:
                        continue
                    M[cols.astype(np.int64)] = (panel.conj() if conj else panel).T @ xin[row0: row0 + h]

    def combine_m(self, s, M, twice_only):
        for ce in self.side[s].combine_m:
            pk = int(ce["packed"])
            n_sum, ln, tw = pk & 0xFFFFFF, (pk >> 24) & 0x7F, pk >> 31
            if twice_only and not tw:
                continue
            src, dst = int(ce["src"]), int(ce["dst_first"])
            M[dst: dst + ln] = M[src: src + n_sum * ln].reshape(n_sum, ln, -1).sum(axis=0)

    def apply_m(self, s, X, in_shift, out, out_shift, M, alpha, beta, twice_only, conj=False):
        """APPLY_M: per RUN, C[rows] += op(panel) B with B row k = scratch vector cols[k] or input row (cols[k] & ~bit31)."""
        side = self.side[s]
        for b in side.order:
            bd = side.blocks[b]
            if twice_only and not (bd["flags"] & 1):
                continue
            acc = np.zeros((int(bd["nrows"]), out.shape[1]), self.dtype)
            for st in range(int(bd["first_stage"]), int(bd["first_stage"] + bd["n_stages"])):
                if twice_only and not (side.stages[st]["flags"] & 1):
                    continue
                for row0, h, twice, panel, cols in side.runs_of_stage(st, self.dtype, "apply"):
                    if twice_only and not twice:
                        continue
                    B = np.zeros((len(cols), X.shape[1]), self.dtype)
                    dense = (cols & 0x80000000) != 0
                    rows = (cols[dense] & 0x7FFFFFFF).astype(np.int64) + in_shift
                    ok = (rows >= 0) & (rows < X.shape[0])
                    Bd = np.zeros((int(dense.sum()), X.shape[1]), self.dtype)
                    Bd[ok] = X[rows[ok]]
                    B[dense] = Bd
                    B[~dense] = M[cols[~dense].astype(np.int64)]
                    acc[row0: row0 + h] += (panel.conj() if conj else panel) @ B
                pairs, n_panel = self._munits_of_stage(s, st)
                for (u, row0, h, w, kind, twice, panel), mu in pairs[n_panel:]:
                    assert kind == UNIT_ADDVEC and (int(mu["flags"]) >> 8) & 0xFF == h
                    if twice_only and not twice:
                        continue
                    src = int(mu["src"])
                    acc[row0: row0 + h] += M[src: src + h]
            for i in range(int(bd["nrows"])):
                g = int(bd["row_start"]) + i + out_shift
                if 0 <= g < out.shape[0]:
                    out[g] = alpha * acc[i] + (0 if beta == 0 else beta * out[g])

    def matrix_product_row_major(self, trans, alpha, X, beta, Y, mu):
        """run_product_m of capi.cu (double and complex<double>): X, Y row-major (n x mu)."""
        sym = self.sym
        if (trans == "T" and sym == "H") or (trans == "C" and sym == "S"):
            return 2
        twice = sym != "N" and self.any_twice
        cplx = self.dtype == np.complex128
        X, Y = X.reshape(-1, mu), Y.reshape(-1, mu)
        nvec = max(1, self.side[0].mscratch_elems)
        M1, M2 = np.full((nvec, mu), np.nan, self.dtype), np.full((nvec, mu), np.nan, self.dtype)
        D = self.D

        def direction(cs, M, in_shift, out_shift, b, twice_only, conj):
            self.reduce_m(1 - cs, X, in_shift, M, twice_only, conj)
            self.combine_m(cs, M, twice_only)
            self.apply_m(cs, X, in_shift, Y, out_shift, M, alpha, b, twice_only, conj)

        herm = sym == "H" and cplx
        if trans == "N":
            direction(0, M1, 0, 0, beta, False, False)
            if twice:
                direction(1, M2, D, -D, 1.0, True, herm)
        else:
            direction(1, M1, 0, 0, beta, False, trans == "C" and cplx)
            if twice:
                direction(0, M2, -D, D, 1.0, True, False)
        return 0

    def vector_product(self, trans, alpha, x, beta, y):
        sym = self.sym
        if (trans == "T" and sym == "H") or (trans == "C" and sym == "S"):
            return 2
        twice = sym != "N" and self.any_twice
        is_complex = self.dtype == np.complex128
        T1, T2 = self.new_scratch(), self.new_scratch()
        D = self.D
        if trans == "N":
            self.reduce(1, x, 0, T1, False, False)
            if twice:
                self.reduce(0, x, D, T2, True, sym == "H" and is_complex)
                self.combine(1, T2, True)
            self.combine(0, T1, False)
            self.apply(0, y, 0, T1, alpha, beta, False, False)
            if twice:
                self.apply(1, y, -D, T2, alpha, 1.0, True, sym == "H" and is_complex)
        else:
            conj = trans == "C" and is_complex
            self.reduce(0, x, 0, T1, False, conj)
            self.combine(1, T1, False)
            if twice:
                self.reduce(1, x, -D, T2, True, False)
                self.combine(0, T2, True)
            self.apply(1, y, 0, T1, alpha, beta, False, conj)
            if twice:
                self.apply(0, y, D, T2, alpha, 1.0, True, False)
        return 0
