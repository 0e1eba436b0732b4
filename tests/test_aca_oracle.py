"""The oracle of the leaf assembly (oracle/aca_oracle.c: the reference's sympartialACA restated in plain C) pinned against
the reference — the fixtures of tests/golden/aca/ (assembled by the unmodified reference, tools/make_golden_aca.py) and the
reference run live (oracle/_ref) — and the packer's low-rank task table (what the device copies out of its factor pool)
interpreted in numpy. CPU only."""
import numpy as np
import pytest
from aca_cases import ACA_GOLDEN, DIAG_FLAGS, AcaCase, packed_side

from htool_b200.capi import LEAF_NP_DTYPE
from oracle.flatcase import oracle_sympartial_aca


def check_all_lowrank_leaves(case):
    f = case.flat
    n_lr = 0
    for i in np.nonzero(f.table[:, 4] >= 0)[0]:
        q, U, V, piv = case.oracle_block(i)
        assert q == f.table[i, 4], (i, q, f.table[i, 4])
        Ur, Vr = case.factors(i)
        assert np.array_equal(U, Ur) and np.array_equal(V, Vr), f"leaf {i}: factors differ from the reference's"
        assert len(np.unique(piv[:, 0])) == q and len(np.unique(piv[:, 1])) == q  # a row / column is never visited twice
        n_lr += 1
    return n_lr


@pytest.mark.parametrize("name", ACA_GOLDEN)
def test_oracle_reproduces_the_reference_factors(name):
    """Same ranks, bit-identical U and V for every low-rank leaf of the reference-assembled fixture."""
    assert check_all_lowrank_leaves(AcaCase.golden(name)) > 0


def test_fixtures_exist():
    assert len(ACA_GOLDEN) >= 8


@pytest.mark.parametrize("kw", [dict(n=2500, kernel="laplace_reg", epsilon=1e-4), dict(n=1800, kernel="laplace_reg", epsilon=1e-5, symmetry="S", uplo="L"),
                                dict(n=1500, n_source=1100, same_cluster=False, z_source=1.5, kernel="laplace", epsilon=1e-3),
                                dict(n=1800, dtype="complex", kernel="complex_reg", epsilon=1e-4, symmetry="S", uplo="L"),
                                dict(n=1600, dtype="complex", kernel="hermitian_reg", epsilon=1e-5, symmetry="H", uplo="U"),
                                dict(n=2000, dtype="complex", kernel="helmholtz", epsilon=1e-4, wavenumber=5.0)], ids=["N", "SL", "rect", "z_SL", "z_HU", "z_helmholtz"])
def test_oracle_against_the_live_reference(kw, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref is not built")
    assert check_all_lowrank_leaves(AcaCase.live(**kw)) > 100


def test_compression_failures():
    """The reference gives up (rank -1: the leaf becomes dense, tree_builder.hpp:619-625) when the next term would cost more
    than the dense block (sympartialACA.hpp:102) — e.g. thin blocks — and on a zero first row."""
    rng = np.random.default_rng(5)
    tp, sp = rng.random((40, 3)), rng.random((40, 3)) + np.array([3.0, 0, 0])
    for m, n in [(1, 1), (1, 7), (5, 1), (2, 2), (3, 2)]:
        q, *_ = oracle_sympartial_aca("laplace", tp, sp, m, n, 100, 0, 0, 0, 1e-12)
        assert q == -1, (m, n, q)
    # a smooth far-field block converges well before the size test bites
    q, U, V, _ = oracle_sympartial_aca("laplace", tp, sp, 40, 40, 100, 0, 0, 0, 1e-6)
    assert 0 < q < 20
    gi, gj = np.meshgrid(np.arange(40), np.arange(40), indexing="ij")
    A = 1.0 / (4 * np.pi * np.linalg.norm(tp[gi] - sp[gj], axis=-1))
    assert np.linalg.norm(A - U @ V) < 1e-5 * np.linalg.norm(A)
    # both orientations (row_offset < col_offset: dimension 1 = the columns) give the block itself, not its transpose
    q2, U2, V2, _ = oracle_sympartial_aca("laplace", tp, sp, 40, 30, 0, 100, 0, 5, 1e-6)
    A2 = 1.0 / (4 * np.pi * np.linalg.norm(tp[np.arange(40)[:, None]] - sp[5 + np.arange(30)[None, :]], axis=-1))
    assert q2 > 0 and np.linalg.norm(A2 - U2 @ V2) < 1e-5 * np.linalg.norm(A2)


def test_symmetric_blocks_are_transposes_of_each_other():
    """sympartialACA.hpp:46-66: block (t, s) with offsets (a, b) and block (s, t) with offsets (b, a) get the same crosses."""
    rng = np.random.default_rng(7)
    pts = rng.random((120, 3)) * np.array([8.0, 1, 1])
    pts = pts[np.argsort(pts[:, 0])]
    q1, U1, V1, _ = oracle_sympartial_aca("laplace_reg", pts, pts, 30, 40, 80, 0, 80, 0, 1e-5)
    q2, U2, V2, _ = oracle_sympartial_aca("laplace_reg", pts, pts, 40, 30, 0, 80, 0, 80, 1e-5)
    assert q1 == q2 > 0 and np.array_equal(U1, V2.T) and np.array_equal(V1, U2.T)


@pytest.mark.parametrize("name", ACA_GOLDEN)
def test_lowrank_task_table_rebuilds_the_host_streams(name):
    """Low-rank leaves without host factors: their panels travel as zeros and the task table tells the device which slice of
    which factor goes where. Filling the panels in numpy from the reference's factors must rebuild the host-packed streams."""
    case = AcaCase.golden(name)
    f = case.flat
    lr = f.table[:, 4] >= 0
    esz = np.dtype(case.dtype).itemsize
    desc0, keep = case.stripped_desc()
    lv = np.frombuffer(keep, dtype=LEAF_NP_DTYPE)
    lv["rank"][: f.table.shape[0]] = f.table[:, 4]  # the ranks are known (after the device ACA), the factors are not on the host
    for side in (0, 1):
        ref_stream, d0, l0 = packed_side(f.desc, side, False)
        assert len(d0) == 0 and len(l0) == 0
        stream, dense, lowrank = packed_side(desc0, side, True)
        assert stream.size == ref_stream.size and len(lowrank) > 0
        assert (lowrank["lcol"] == side).all() and lr[lowrank["lrow"]].all()
        if side == 1:
            assert len(dense) == 0
        covered = np.zeros(f.table.shape[0], np.int64)
        for t in lowrank:
            i = int(t["lrow"])
            U, V = case.factors(i)
            h, w, ld, p0, k0 = int(t["h"]), int(t["w"]), int(t["ld"]), int(t["p0"]), int(t["k0"])
            src = U[p0: p0 + h, k0: k0 + w] if side == 0 else V[k0: k0 + w, p0: p0 + h].T
            panel = np.zeros((w, ld), case.dtype)
            panel[:, :h] = src.T
            off = int(t["byte_off"])
            stream[off: off + w * ld * esz] = np.frombuffer(panel.tobytes(), dtype=np.uint8)
            covered[i] += h * w
        m, n, r = f.table[:, 2].astype(np.int64), f.table[:, 3].astype(np.int64), f.table[:, 4].astype(np.int64)
        assert np.array_equal(covered[lr], ((m if side == 0 else n) * r)[lr]), "every factor entry is copied exactly once"
        if side == 0:  # the dense panels are the business of tests/test_generated_dense.py: take them from the host stream
            for t in dense:
                off, nbytes = int(t["byte_off"]), int(t["w"]) * int(t["ld"]) * esz
                stream[off: off + nbytes] = ref_stream[off: off + nbytes]
        assert np.array_equal(stream, ref_stream)


def test_diag_flags_constant():
    assert DIAG_FLAGS == 0x6


@pytest.mark.parametrize("name", ["d_SL", "d_rect", "z_HU"])
def test_headers_only_upload_rebuilds_the_stream(name):
    """Device assembly (no leaf carries host data): only the stage headers travel (htb_pack_host: headers + header_offsets) and a
    kernel puts them into the zeroed stream. Zeros + the scattered headers must equal the stream the packer fills on the host."""
    import ctypes as C

    from aca_cases import TASK_DT  # noqa: F401  (same module: keeps the struct layouts in one place)
    from htool_b200 import capi

    case = AcaCase.golden(name)
    f = case.flat
    desc0, keep = case.stripped_desc()
    lv = np.frombuffer(keep, dtype=LEAF_NP_DTYPE)
    lv["rank"][: f.table.shape[0]] = f.table[:, 4]  # (ranks known, no data anywhere)
    lib = capi.load()
    capi.set_option("pack_generate_dense", 1)
    try:
        for side in (0, 1):
            p = capi.htb_packed_side()
            capi.check(lib, lib.htb_pack_host(C.byref(desc0), side, C.byref(p)))
            assert p.header_bytes > 0 and p.headers and p.header_offsets
            stream = np.frombuffer((C.c_char * p.stream_bytes).from_address(p.stream), dtype=np.uint8).copy()
            headers = np.frombuffer((C.c_char * p.header_bytes).from_address(p.headers), dtype=np.uint8)
            offs = np.frombuffer((C.c_char * (8 * (p.n_stages + 1))).from_address(p.header_offsets), dtype=np.uint64).astype(np.int64)
            stages = np.frombuffer((C.c_char * (32 * p.n_stages)).from_address(p.stages), dtype=np.dtype([("byte_off", "<u8"), ("rest", "V24")]))
            rebuilt = np.zeros_like(stream)
            for st in range(p.n_stages):
                n = int(offs[st + 1] - offs[st])
                o = int(stages["byte_off"][st])
                rebuilt[o: o + n] = headers[int(offs[st]): int(offs[st]) + n]
            assert int(offs[-1]) == p.header_bytes and p.header_bytes < 0.25 * p.stream_bytes
            assert np.array_equal(rebuilt, stream)
            lib.htb_pack_free(C.byref(p))
        # a descriptor whose leaves carry data has no header-only form
        p = capi.htb_packed_side()
        capi.check(lib, lib.htb_pack_host(C.byref(f.desc), 0, C.byref(p)))
        assert p.header_bytes == 0 and not p.headers
        lib.htb_pack_free(C.byref(p))
    finally:
        capi.set_option("pack_generate_dense", 0)
