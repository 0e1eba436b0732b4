"""CPU tests of the host logic: C-ABI surface, packer / stream format and pass sequences.

The packed bytes (what htb_create would upload) are obtained through htb_pack_host and interpreted by
tests/stream_emulator.py with the same walk as the CUDA kernels; results are compared with the oracle and
the golden reference vectors. No compute entry point is called (there is no GPU here and no CPU fallback:
htb_create must fail with HTB_ERR_CUDA).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest
from conftest import GOLDEN, REPO, load_golden, rel_err, rnd, valid_trans

from htool_b200 import capi
from oracle.flatcase import random_flatcase
from stream_emulator import Emulator, PackedSide

TOL = 1e-13


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    header = open(os.path.join(REPO, "include", "htool_b200.h")).read()
    declared = set(re.findall(r"\b(htb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    assert C.sizeof(capi.htb_leaf) == 40
    assert C.sizeof(capi.htb_hmatrix_desc) == 48
    assert C.sizeof(capi.htb_info) == 8 * 9 + 4 * 10


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    flat = random_flatcase(seed=0)
    with pytest.raises(capi.HtbError) as ei:
        capi.Operator(flat.desc)
    assert ei.value.status == capi.HTB_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_invalid_descriptions_are_rejected():
    lib = capi.load()
    flat = random_flatcase(seed=1, n_leaves=5)
    bad = random_flatcase(seed=1, n_leaves=5)
    bad.table[2, 0] = bad.nb_rows  # leaf outside the root block
    bad._build_desc()
    p = capi.htb_packed_side()
    assert lib.htb_pack_host(C.byref(bad.desc), 0, C.byref(p)) == capi.HTB_ERR_INVALID
    assert b"outside" in lib.htb_last_error()
    assert lib.htb_pack_host(C.byref(flat.desc), 2, C.byref(p)) == capi.HTB_ERR_INVALID
    assert lib.htb_set_option(b"no_such_option", 1) == capi.HTB_ERR_INVALID


def _check(flat, tol=TOL):
    em = Emulator(flat)
    rng = np.random.default_rng(3)
    for trans in "NTC":
        ni, no = (flat.nb_cols, flat.nb_rows) if trans == "N" else (flat.nb_rows, flat.nb_cols)
        for alpha, beta in [(1.0, 0.0), (0.7, -1.3)]:
            if flat.np_dtype == np.complex128:
                alpha, beta = alpha * (1 + 0.5j), beta * (1 - 0.25j)
            x, y0 = rnd(rng, ni, flat.np_dtype), rnd(rng, no, flat.np_dtype)
            yo, ye = y0.copy(), y0.copy()
            st = flat.oracle_vector_product(trans, alpha, x, beta, yo)
            assert em.vector_product(trans, alpha, x, beta, ye) == st
            if st == 0:
                assert rel_err(ye, yo) < tol, (trans, alpha, beta)
    return em


@pytest.mark.parametrize("seed", range(3))
@pytest.mark.parametrize("dtype_code,symmetric", [(0, None), (1, None), (0, "S"), (1, "S"), (1, "H")])
def test_packed_stream_random_leaves(seed, dtype_code, symmetric):
    _check(random_flatcase(seed=seed, dtype_code=dtype_code, symmetric=symmetric))


@pytest.mark.parametrize("name", GOLDEN)
def test_packed_stream_matches_golden(name):
    flat, entries, _ = load_golden(name)
    em = Emulator(flat)
    for e in entries:
        if e["mu"] != 1:
            continue
        y = e["y_in"].copy()
        assert em.vector_product(e["trans"], e["alpha"], e["x"], e["beta"], y) == 0
        assert rel_err(y, e["y_seq"]) < TOL, (name, e["trans"])


@pytest.mark.parametrize("opts", [dict(block_rows=32, piece_cols=4, stage_bytes=4096, cseg_bytes=512), dict(block_rows=128, piece_cols=32, stage_bytes=65536, cseg_bytes=4096),
                                  dict(piece_cols=8, stage_bytes=16384, cseg_bytes=2048)])
def test_packer_options(opts):
    DEFAULTS = {k: capi.get_option(k) for k in ("block_rows", "piece_cols", "stage_bytes", "cseg_bytes")}
    try:
        for k, v in opts.items():
            capi.set_option(k, v)
        _check(random_flatcase(seed=5, symmetric="S"))
        if opts.get("block_rows", 64) * 16 <= 1024:
            _check(random_flatcase(seed=6, dtype_code=1))
        else:  # a lane owns at most two row slabs: block_rows * sizeof(T) <= 1024
            with pytest.raises(capi.HtbError):
                PackedSide(random_flatcase(seed=6, dtype_code=1).desc, 0)
    finally:
        for k, v in DEFAULTS.items():
            capi.set_option(k, v)


def test_target_side_block_height():
    """Side 0 may be cut into smaller blocks than side 1 (store.hpp: PackOptions.target_block_rows); products stay the same."""
    flat = random_flatcase(seed=8, nb_rows=900, nb_cols=700, n_leaves=200, max_dim=400)
    try:
        assert PackedSide(flat.desc, 0).blocks["nrows"].max() > 64
        for tb in (64, 32):
            capi.set_option("target_block_rows", tb)
            assert PackedSide(flat.desc, 0).blocks["nrows"].max() <= tb < PackedSide(flat.desc, 1).blocks["nrows"].max()
            _check(flat)
    finally:
        capi.set_option("target_block_rows", 0)


def test_tail_split_of_target_blocks():
    """Side 0 of a shard with a partial last round of APPLY CTAs: the rows of that round are re-cut into quarter-height
    blocks (packer.cpp: make_blocks); the product does not change."""
    flat = random_flatcase(seed=9, nb_rows=1000, nb_cols=700, n_leaves=220, max_dim=400)
    try:
        capi.set_option("tail_split", 0)
        plain = PackedSide(flat.desc, 0)
        nb = plain.n_blocks
        slots = next(sl for sl in range(3, nb) if 0 < nb % sl <= 0.8 * sl and nb <= 6 * sl)
        capi.set_option("tail_split", 1)
        capi.set_option("cta_slots", slots)
        split = PackedSide(flat.desc, 0)
        r = nb % slots
        assert split.n_blocks > nb
        assert np.array_equal(split.blocks["row_start"][: nb - r], plain.blocks["row_start"][: nb - r])  # full rounds untouched
        assert split.blocks["nrows"][nb - r:].max() <= 32
        assert PackedSide(flat.desc, 1).n_blocks == (lambda: (capi.set_option("tail_split", 0), PackedSide(flat.desc, 1).n_blocks, capi.set_option("tail_split", 1))[1])()
        _check(flat)
    finally:
        capi.set_option("tail_split", 1)
        capi.set_option("cta_slots", 0)


def test_stream_invariants():
    flat, _, _ = load_golden("d_N")
    for s in (0, 1):
        side = PackedSide(flat.desc, s)
        # blocks tile the index space, in order
        assert side.blocks["row_start"][0] == 0
        assert (side.blocks["row_start"][1:] == side.blocks["row_start"][:-1] + side.blocks["nrows"][:-1]).all()
        assert side.blocks["row_start"][-1] + side.blocks["nrows"][-1] == side.n
        assert (side.blocks["nrows"] <= 128).all() and (side.blocks["nrows"] > 0).all()
        # stages are 16-byte aligned, contiguous and within the granule
        assert (side.stages["byte_off"] % 16 == 0).all() and (side.stages["nbytes"] % 16 == 0).all()
        assert (side.stages["nbytes"] <= capi.get_option("stage_bytes")).all()
        assert (side.stages["byte_off"][1:] == side.stages["byte_off"][:-1] + side.stages["nbytes"][:-1]).all()
        assert sorted(side.order.tolist()) == list(range(side.n_blocks))
    # every coefficient is stored exactly once per side it is needed on: U + dense on side 0, V on side 1
    tbl = flat.table
    dense = tbl[:, 4] < 0
    side0 = int((tbl[dense, 2].astype(np.int64) * tbl[dense, 3]).sum() + (tbl[~dense, 2].astype(np.int64) * tbl[~dense, 4]).sum())
    side1 = int((tbl[~dense, 3].astype(np.int64) * tbl[~dense, 4]).sum())
    for s, expect in ((0, side0), (1, side1)):
        side = PackedSide(flat.desc, s)
        n = 0
        for st in range(len(side.stages)):
            for u, row0, h, w, kind, twice, panel in side.units_of_stage(st, flat.np_dtype):
                n += 0 if panel is None else h * w
        assert n == expect


def test_empty_and_degenerate_operators():
    from oracle.flatcase import FlatCase

    # no leaves at all: out <- beta * out
    f = FlatCase(0, 37, 21, 0, 0, "N", "N", np.zeros((0, 6), np.int32), np.zeros(0))
    em = Emulator(f)
    x, y = np.ones(21), np.full(37, 2.0)
    assert em.vector_product("N", 1.0, x, 0.5, y) == 0
    assert np.allclose(y, 1.0)
    # rank-0 low-rank leaf and 1 x 1 leaves
    tbl = np.array([[0, 0, 5, 4, 0, 0], [3, 2, 1, 1, -1, 0], [1, 1, 1, 1, 1, 0]], np.int32)
    f = FlatCase(0, 6, 5, 0, 0, "N", "N", tbl, np.array([2.0, 3.0, 4.0]))
    _check(f)


@pytest.mark.parametrize("seed", range(2))
@pytest.mark.parametrize("symmetric,dtype", [(None, 0), ("S", 0), (None, 1), ("S", 1), ("H", 1)])
def test_multi_rhs_side_tables(seed, symmetric, dtype):
    """RUNS + column tables (aux records), TF / PARTM layout and MUnit tables of the multi-RHS path (mkernels.cu), emulated,
    for double and complex<double>."""
    flat = random_flatcase(seed=seed, symmetric=symmetric, dtype_code=dtype)
    em = Emulator(flat)
    rng = np.random.default_rng(1)
    dt = flat.np_dtype
    for trans in valid_trans(flat.symmetry):
        ni, no = (flat.nb_cols, flat.nb_rows) if trans == "N" else (flat.nb_rows, flat.nb_cols)
        mu = 5
        a, b = (0.5, 2.0) if dt == np.float64 else (0.5 - 0.3j, 2.0 + 0.25j)
        X, Y0 = rnd(rng, ni * mu, dt), rnd(rng, no * mu, dt)
        Yo, Ye = Y0.copy(), Y0.copy()
        assert flat.oracle_matrix_product_row_major(trans, a, X, b, Yo, mu) == 0
        assert em.matrix_product_row_major(trans, a, X, b, Ye, mu) == 0
        assert rel_err(Ye, Yo) < TOL, trans


def test_runs_group_the_units_of_a_cluster():
    """Inside a block the packer orders the units by the rows they act on: a stage's runs are few and wide."""
    flat, _, _ = load_golden("d_N")
    em = Emulator(flat)
    for s in range(2):
        side = em.side[s]
        n_runs = n_units = 0
        for st in range(len(side.stages)):
            n_runs += sum(1 for _ in side.runs_of_stage(st, em.dtype, "apply"))
            n_units += int(side.stages[st]["n_panel"])
        assert n_runs * 2 <= n_units, (s, n_runs, n_units)


def test_multi_rhs_side_tables_golden():
    for name in ("d_N", "d_SL", "d_strip_SU", "d_rect", "z_HL", "z_SL", "z_N_helmholtz"):
        flat, entries, _ = load_golden(name)
        em = Emulator(flat)
        for e in entries:
            if e["mu"] == 1:
                continue
            y = e["y_in"].copy()
            assert em.matrix_product_row_major(e["trans"], e["alpha"], e["x"], e["beta"], y, e["mu"]) == 0
            assert rel_err(y, e["y_seq"]) < TOL, (name, e["trans"])


@pytest.mark.parametrize("dtype_code,symmetric", [(0, "-"), (1, "H")])
def test_layout_does_not_depend_on_the_thread_count(dtype_code, symmetric):
    """The packer's serial passes run one side per thread and its block loops are dynamically scheduled: streams and tables must
    be the same bytes whatever OMP_NUM_THREADS is (1 = the serial fallback of both_sides, packer.cpp)."""
    import subprocess
    import sys

    digests = set()
    for threads in ("1", "2", "7"):
        r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "pack_digest.py"), str(dtype_code), symmetric], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, OMP_NUM_THREADS=threads))
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
        digests.add([l for l in r.stdout.splitlines() if l.startswith("digest ")][-1])
    assert len(digests) == 1, digests
