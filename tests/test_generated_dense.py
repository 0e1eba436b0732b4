"""Leaf assembly on the device, first step: the dense leaves of an H-matrix generated on the GPU
straight into the leaf store (htb_create_generated, htool_b200/csrc/generate.cu) instead of
HMatrix::compute_dense_data on the host (hmatrix.hpp:222-226).

CPU (`-m "not gpu"`): the packer's task table is interpreted in numpy — every task's coefficients, evaluated with the
reference's kernel formulas at the points, must rebuild the stream htb_pack_host produces from the reference's dense data.
GPU (`-m gpu`): the device-generated store is downloaded and compared with the host-packed one: identical descriptors
everywhere, BIT-IDENTICAL coefficients for the real kernels, <= 1e-15 relative (to |z|) for Helmholtz (sin / cos); the
products agree with the reference's.
"""
import ctypes as C

import numpy as np
import pytest
from conftest import rel_err

from htool_b200 import capi

TASK_DT = np.dtype([("byte_off", "<u8"), ("lrow", "<i4"), ("lcol", "<i4"), ("p0", "<i4"), ("k0", "<i4"), ("h", "<u2"), ("w", "<u2"), ("ld", "<u2"), ("flags", "<u2")])
assert TASK_DT.itemsize == 32

CASES = [
    ("laplace_reg", dict(n=1500)),
    ("laplace_reg", dict(n=1200, symmetry="S", uplo="L")),
    ("laplace_reg", dict(n=1200, symmetry="S", uplo="U")),
    ("laplace", dict(n=900, n_source=700, same_cluster=False, z_source=1.5, kernel="laplace")),
    ("complex_reg", dict(n=1000, dtype="complex", kernel="complex_reg", symmetry="S", uplo="L")),
    ("hermitian_reg", dict(n=1000, dtype="complex", kernel="hermitian_reg", symmetry="H", uplo="U")),
    ("helmholtz", dict(n=1300, dtype="complex", kernel="helmholtz", wavenumber=5.0)),
]


def kernel_values(kernel, a, b, k):
    """generator_test.hpp:155-205 / ref_harness.hpp KernelGenerator, same operation order (numpy float64 = IEEE double)."""
    dx, dy, dz = a[..., 0] - b[..., 0], a[..., 1] - b[..., 1], a[..., 2] - b[..., 2]
    r = np.sqrt((dx * dx + dy * dy) + dz * dz)
    fpr = (4 * np.pi) * r
    with np.errstate(divide="ignore", invalid="ignore"):
        if kernel == "laplace":
            return 1.0 / fpr
        if kernel == "laplace_reg":
            return 1.0 / (1e-5 + fpr)
        if kernel == "complex":
            v = 1.0 / fpr
            return v + 1j * v
        if kernel == "complex_reg":
            v = 1.0 / (1e-5 + fpr)
            return v + 1j * v
        if kernel == "hermitian_reg":
            d = 1e-5 + fpr
            return 1.0 / d + 1j * (np.sign(dx) / d)
        z = (np.cos(k * r) + 1j * np.sin(k * r)) / fpr
        return np.where(r < 1e-12, 1.0 / ((4 * np.pi) * 1e-3) + 1j * (k / (4 * np.pi)), z)


def packed(desc, generate):
    lib = capi.load()
    capi.set_option("pack_generate_dense", 1 if generate else 0)
    try:
        p = capi.htb_packed_side()
        capi.check(lib, lib.htb_pack_host(C.byref(desc), 0, C.byref(p)))
        stream = np.frombuffer((C.c_char * p.stream_bytes).from_address(p.stream), dtype=np.uint8).copy() if p.stream_bytes else np.zeros(0, np.uint8)
        tasks = np.frombuffer((C.c_char * (p.n_dense_tasks * 32)).from_address(p.dense_tasks), dtype=TASK_DT).copy() if p.n_dense_tasks else np.zeros(0, TASK_DT)
        lib.htb_pack_free(C.byref(p))
    finally:
        capi.set_option("pack_generate_dense", 0)
    return stream, tasks


def make_case(kw):
    from oracle import refharness as R

    if not R.available():
        pytest.skip("oracle/_ref is not built")
    return R.RefCase(**kw)


@pytest.mark.parametrize("kernel,kw", CASES, ids=[f"{c[0]}_{i}" for i, c in enumerate(CASES)])
def test_task_table_rebuilds_the_host_stream(kernel, kw):
    case = make_case(kw)
    ref_stream, no_tasks = packed(case.desc, False)
    assert len(no_tasks) == 0
    desc0, keep = case.desc_without_dense_data()
    stream, tasks = packed(desc0, True)
    lv = case.leaves()
    n_dense_units_min = int((lv["rank"] < 0).sum())
    assert len(tasks) >= n_dense_units_min > 0 and stream.size == ref_stream.size
    assert not np.array_equal(stream, ref_stream)  # the dense panels travel as zeros
    tp, sp = case.points(0), case.points(1)
    dt = case.np_dtype
    esz = np.dtype(dt).itemsize
    k = float(kw.get("wavenumber", 0.0))
    for t in tasks:
        h, w, ld = int(t["h"]), int(t["w"]), int(t["ld"])
        gi = int(t["p0"]) + np.arange(h)[:, None] + np.zeros((1, w), np.int64)
        gj = int(t["k0"]) + np.arange(w)[None, :] + np.zeros((h, 1), np.int64)
        flags = int(t["flags"])
        mirrored = np.zeros((h, w), bool)
        if flags & 0x6:  # symv / hemv leaf: only the UPLO triangle is read
            stored = (gi <= gj) if (flags & 0x8) else (gi >= gj)
            mirrored = ~stored
            gi, gj = np.where(mirrored, gj, gi), np.where(mirrored, gi, gj)
        v = kernel_values(kernel, tp[int(t["lrow"]) + gi], sp[int(t["lcol"]) + gj], k).astype(dt)
        if flags & 0x4:
            v = np.where(mirrored, np.conj(v), v)
            v = np.where(gi == gj, v.real, v)
        panel = np.zeros((w, ld), dt)
        panel[:, :h] = v.T
        off = int(t["byte_off"])
        stream[off: off + w * ld * esz] = np.frombuffer(panel.tobytes(), dtype=np.uint8)
    if kernel == "helmholtz":
        a, b = stream.view(np.float64), ref_stream.view(np.float64)
        same = a == b
        assert same.mean() > 0.9 and np.allclose(a[~same], b[~same], rtol=1e-13, atol=1e-300)
    else:
        assert np.array_equal(stream, ref_stream), "the task table + the kernel formulas must rebuild the host-packed stream bit for bit"


@pytest.mark.gpu
@pytest.mark.parametrize("kernel,kw", CASES, ids=[f"{c[0]}_{i}" for i, c in enumerate(CASES)])
def test_device_generated_store_matches_compute_dense_data(kernel, kw):
    case = make_case(kw)
    ref_stream, _ = packed(case.desc, False)
    desc0, keep = case.desc_without_dense_data()
    k = float(kw.get("wavenumber", 0.0))
    op = capi.Operator(desc0, generator=(kernel, case.points(0), case.points(1), k))
    got = op.download_store(0, ref_stream.size)
    if kernel == "helmholtz":
        a, b = got.view(np.complex128), ref_stream.view(np.complex128)
        # descriptors are integers: compare them as bytes where the streams differ as numbers
        diff = np.abs(a - b)
        scale = np.maximum(np.abs(b), 1e-300)
        ok = (got.view(np.uint64) == ref_stream.view(np.uint64)).reshape(-1, 2).all(axis=1) | (diff <= 4e-15 * scale)
        assert ok.all(), float((diff / scale)[~ok].max())
    else:
        assert np.array_equal(got, ref_stream), "device-generated dense leaves differ from compute_dense_data"
    # and the products: reference CPU product on the host-generated H-matrix vs the device-generated operator
    rng = np.random.default_rng(3)
    dt = case.np_dtype
    x = (rng.random(case.nb_cols) - 0.5).astype(dt)
    if dt == np.complex128:
        x = x + 1j * (rng.random(case.nb_cols) - 0.5)
    y_ref, y = np.zeros(case.nb_rows, dt), np.zeros(case.nb_rows, dt)
    case.vector_product("N", 1.0, x, 0.0, y_ref, variant="openmp")
    op.add_vector_product("N", 1.0, x, 0.0, y)
    assert rel_err(y, y_ref) < 1e-12
    # a second operator from the host data: the two stores give the same bits for the real kernels
    op2 = capi.Operator(case.desc)
    y2 = np.zeros(case.nb_rows, dt)
    op2.add_vector_product("N", 1.0, x, 0.0, y2)
    if kernel != "helmholtz":
        assert np.array_equal(y, y2)
    op.close()
    op2.close()


@pytest.mark.gpu
def test_generator_arguments_are_validated():
    case = make_case(dict(n=300))
    desc0, keep = case.desc_without_dense_data()
    with pytest.raises(capi.HtbError) as ei:  # plain htb_create refuses leaves without data
        capi.Operator(desc0)
    assert ei.value.status == capi.HTB_ERR_INVALID
    with pytest.raises(capi.HtbError) as ei:  # a complex kernel cannot fill a double H-matrix
        capi.Operator(desc0, generator=("helmholtz", case.points(0), case.points(1), 1.0))
    assert ei.value.status == capi.HTB_ERR_INVALID
