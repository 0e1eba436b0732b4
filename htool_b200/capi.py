"""ctypes view of include/htool_b200.h — used by tests/ and bench.py only.

The product is the C-ABI shared library (htool_b200/lib/libhtool_b200.so) plus the header-only C++ shim in
htool_b200/cpp/; Python is not part of it. This module mirrors the structs of the header 1:1 so the parity
tests and the benchmark can call the same entry points the C++ shim calls.

There is no fallback: if the library is missing, `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libhtool_b200.so")

HTB_OK, HTB_ERR_INVALID, HTB_ERR_UNSUPPORTED, HTB_ERR_CUDA, HTB_ERR_NCCL, HTB_ERR_ALLOC = range(6)
HTB_DOUBLE, HTB_COMPLEX_DOUBLE = 0, 1
HTB_MEM_HOST, HTB_MEM_DEVICE = 0, 1
HTB_LEAF_APPLY_TRANSPOSED_TOO = 0x1
HTB_LEAF_DIAG_SYMMETRIC = 0x2
HTB_LEAF_DIAG_HERMITIAN = 0x4
HTB_LEAF_UPLO_UPPER = 0x8
HTB_NCCL_UNIQUE_ID_BYTES = 128


class htb_leaf(C.Structure):
    _fields_ = [
        ("row_offset", C.c_int32),
        ("col_offset", C.c_int32),
        ("nb_rows", C.c_int32),
        ("nb_cols", C.c_int32),
        ("rank", C.c_int32),
        ("flags", C.c_int32),
        ("data0", C.c_void_p),
        ("data1", C.c_void_p),
    ]


class htb_hmatrix_desc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("nb_rows", C.c_int32),
        ("nb_cols", C.c_int32),
        ("row_offset", C.c_int32),
        ("col_offset", C.c_int32),
        ("symmetry_for_leaves", C.c_char),
        ("uplo_for_leaves", C.c_char),
        ("reserved", C.c_char * 2),
        ("device", C.c_int32),
        ("nb_leaves", C.c_int64),
        ("leaves", C.POINTER(htb_leaf)),
    ]


class htb_info(C.Structure):
    _fields_ = [
        ("nb_leaves", C.c_int64),
        ("nb_dense_leaves", C.c_int64),
        ("nb_low_rank_leaves", C.c_int64),
        ("nb_leaves_applied_twice", C.c_int64),
        ("coefficients", C.c_int64),
        ("coefficients_twice", C.c_int64),
        ("store_bytes", C.c_int64),
        ("descriptor_bytes", C.c_int64),
        ("workspace_bytes", C.c_int64),
        ("rank_min", C.c_int32),
        ("rank_max", C.c_int32),
        ("dtype", C.c_int32),
        ("device", C.c_int32),
        ("nb_rows", C.c_int32),
        ("nb_cols", C.c_int32),
        ("nb_target_blocks", C.c_int32),
        ("nb_source_blocks", C.c_int32),
        ("sm_count", C.c_int32),
        ("dist_gather", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class htb_packed_side(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("n_blocks", C.c_int32),
        ("n_stages", C.c_int64),
        ("n_combine", C.c_int64),
        ("n_combine_dst", C.c_int64),
        ("stream_bytes", C.c_int64),
        ("scratch_elems", C.c_int64),
        ("cs_base", C.c_int64),
        ("cs_elems", C.c_int64),
        ("part_base", C.c_int64),
        ("part_elems", C.c_int64),
        ("piece_cols", C.c_int32),
        ("block_rows", C.c_int32),
        ("n_munits", C.c_int64),
        ("n_combine_m", C.c_int64),
        ("mscratch_elems", C.c_int64),
        ("blocks", C.c_void_p),
        ("stages", C.c_void_p),
        ("order", C.c_void_p),
        ("combine", C.c_void_p),
        ("combine_dst", C.c_void_p),
        ("stream", C.c_void_p),
        ("munits", C.c_void_p),
        ("combine_m", C.c_void_p),
        ("owner", C.c_void_p),
        ("aux_bytes", C.c_int64),
        ("aux_reduce", C.c_void_p),
        ("aux_apply", C.c_void_p),
        ("n_dense_tasks", C.c_int64),
        ("dense_tasks", C.c_void_p),
        ("n_lowrank_tasks", C.c_int64),
        ("lowrank_tasks", C.c_void_p),
        ("header_bytes", C.c_int64),
        ("headers", C.c_void_p),
        ("header_offsets", C.c_void_p),
    ]


class htb_generator_desc(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("spatial_dimension", C.c_int32), ("wavenumber", C.c_double), ("target_points", C.c_void_p), ("source_points", C.c_void_p)]


class htb_compression_info(C.Structure):
    _fields_ = [("nb_blocks", C.c_int64), ("nb_failed", C.c_int64), ("coefficients", C.c_int64), ("pool_bytes", C.c_int64), ("rank_min", C.c_int32), ("rank_max", C.c_int32),
                ("seconds_aca", C.c_double), ("seconds_total", C.c_double), ("seconds_aca_team", C.c_double * 3), ("nb_blocks_team", C.c_int64 * 3),
                ("seconds_layout", C.c_double), ("seconds_upload", C.c_double), ("seconds_fill", C.c_double), ("seconds_prepare", C.c_double),
                ("seconds_compress", C.c_double), ("seconds_store", C.c_double)]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if hasattr(getattr(self, k), "__len__") else getattr(self, k)) for k, _ in self._fields_}


HTB_RANK_COMPRESS = -2
HTB_KERNELS = {"laplace": 0, "laplace_reg": 1, "complex_reg": 2, "hermitian_reg": 3, "helmholtz": 4, "complex": 5}


class htb_gmres_options(C.Structure):
    _fields_ = [("restart", C.c_int32), ("max_iterations", C.c_int32), ("tolerance", C.c_double), ("orthogonalization", C.c_int32), ("verbosity", C.c_int32),
                ("compute_true_residual", C.c_int32), ("reserved", C.c_int32)]


class htb_gmres_result(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("matvecs", C.c_int32), ("reserved", C.c_int32), ("relative_residual", C.c_double),
                ("true_relative_residual", C.c_double)]


HTB_GMRES_CGS, HTB_GMRES_CGS2 = 0, 1

# name -> (restype, argtypes): every symbol include/htool_b200.h declares
SYMBOLS = {
    "htb_create": (C.c_int, [C.POINTER(htb_hmatrix_desc), C.POINTER(C.c_void_p)]),
    "htb_create_generated": (C.c_int, [C.POINTER(htb_hmatrix_desc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "htb_create_compressed": (C.c_int, [C.POINTER(htb_hmatrix_desc), C.c_void_p, C.c_double, C.POINTER(C.c_void_p)]),
    "htb_get_leaf_ranks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "htb_get_compression_info": (C.c_int, [C.c_void_p, C.POINTER(htb_compression_info)]),
    "htb_download_store": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
    "htb_destroy": (C.c_int, [C.c_void_p]),
    "htb_get_info": (C.c_int, [C.c_void_p, C.POINTER(htb_info)]),
    "htb_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "htb_synchronize": (C.c_int, [C.c_void_p]),
    "htb_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "htb_host_unregister": (C.c_int, [C.c_void_p]),
    "htb_add_vector_product": (C.c_int, [C.c_void_p, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "htb_add_matrix_product_row_major": (C.c_int, [C.c_void_p, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "htb_set_permutations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "htb_add_vector_product_user_numbering": (C.c_int, [C.c_void_p, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "htb_add_matrix_product_user_numbering": (C.c_int, [C.c_void_p, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "htb_nccl_get_unique_id": (C.c_int, [C.c_void_p]),
    "htb_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "htb_comm_destroy": (C.c_int, [C.c_void_p]),
    "htb_dist_add_product_local_to_local": (C.c_int, [C.c_void_p, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "htb_dist_add_product_global_to_global": (C.c_int, [C.c_void_p, C.c_char, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "htb_gmres_default_options": (C.c_int, [C.POINTER(htb_gmres_options)]),
    "htb_gmres": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(htb_gmres_options), C.POINTER(htb_gmres_result), C.c_int]),
    "htb_last_error": (C.c_char_p, []),
    "htb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "htb_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "htb_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "htb_get_option": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64)]),
    "htb_profile_passes": (C.c_int, [C.c_void_p, C.c_int]),
    "htb_get_pass_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "htb_pack_host": (C.c_int, [C.POINTER(htb_hmatrix_desc), C.c_int, C.POINTER(htb_packed_side)]),
    "htb_pack_free": (C.c_int, [C.POINTER(htb_packed_side)]),
}

_lib = None


def load(path: str | None = None):
    """Loads libhtool_b200.so and types every entry point. Raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class HtbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"htool_b200 status {status}: {message}")
        self.status = status


def check(lib, status):
    if status != HTB_OK:
        raise HtbError(status, (lib.htb_last_error() or b"").decode())


def np_dtype(dtype_code: int):
    return np.float64 if dtype_code == HTB_DOUBLE else np.complex128


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class Operator:
    """Thin RAII wrapper over an htb_handle (what htool_b200::GpuHMatrix is on the C++ side)."""

    def __init__(self, desc: htb_hmatrix_desc, keepalive=None, generator=None, compress_epsilon=None):
        """generator = (kernel name, target points (n x 3, cluster numbering), source points, wavenumber): dense leaves
        whose data0 is NULL are generated on the device (htb_create_generated). compress_epsilon: leaves of rank
        HTB_RANK_COMPRESS are compressed on the device as well (htb_create_compressed)."""
        self.lib = load()
        self.handle = C.c_void_p()
        self.dtype_code = desc.dtype
        self.dtype = np_dtype(desc.dtype)
        self.nb_rows, self.nb_cols = desc.nb_rows, desc.nb_cols
        self._keepalive = keepalive
        if generator is None:
            check(self.lib, self.lib.htb_create(C.byref(desc), C.byref(self.handle)))
        else:
            kernel, tp, sp, k = generator
            tp, sp = np.ascontiguousarray(tp, dtype=np.float64), np.ascontiguousarray(sp, dtype=np.float64)
            assert tp.shape == (desc.nb_rows, 3) and sp.shape == (desc.nb_cols, 3)
            g = htb_generator_desc(HTB_KERNELS[kernel], 3, float(k), tp.ctypes.data, sp.ctypes.data)
            if compress_epsilon is None:
                check(self.lib, self.lib.htb_create_generated(C.byref(desc), C.byref(g), C.byref(self.handle)))
            else:
                check(self.lib, self.lib.htb_create_compressed(C.byref(desc), C.byref(g), float(compress_epsilon), C.byref(self.handle)))
        self.nb_leaves = desc.nb_leaves

    def leaf_ranks(self) -> np.ndarray:
        out = np.zeros(self.nb_leaves, dtype=np.int32)
        check(self.lib, self.lib.htb_get_leaf_ranks(self.handle, _ptr(out), self.nb_leaves))
        return out

    def compression_info(self) -> dict:
        i = htb_compression_info()
        check(self.lib, self.lib.htb_get_compression_info(self.handle, C.byref(i)))
        return i.as_dict()

    def download_store(self, side: int, nbytes: int) -> np.ndarray:
        out = np.zeros(nbytes, dtype=np.uint8)
        check(self.lib, self.lib.htb_download_store(self.handle, side, _ptr(out), nbytes))
        return out

    def close(self):
        if self.handle:
            self.lib.htb_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        i = htb_info()
        check(self.lib, self.lib.htb_get_info(self.handle, C.byref(i)))
        return i.as_dict()

    def _scalar(self, v):
        return np.array([v], dtype=self.dtype)

    def set_stream(self, cuda_stream: int | None):
        check(self.lib, self.lib.htb_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        check(self.lib, self.lib.htb_synchronize(self.handle))

    def launch_count(self) -> int:
        n = C.c_int64()
        check(self.lib, self.lib.htb_launch_count(self.handle, C.byref(n)))
        return n.value

    def profile_passes(self, enable: bool):
        check(self.lib, self.lib.htb_profile_passes(self.handle, 1 if enable else 0))

    def pass_times(self) -> dict:
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        check(self.lib, self.lib.htb_get_pass_times(self.handle, ms, n))
        return {k: {"ms": ms[i], "launches": n[i]} for i, k in enumerate(("reduce", "combine", "apply", "other"))}

    # host numpy arrays (HTB_MEM_HOST) -----------------------------------------------------------
    def add_vector_product(self, trans, alpha, x, beta, y):
        a, b = self._scalar(alpha), self._scalar(beta)
        assert x.dtype == self.dtype and y.dtype == self.dtype and x.flags.c_contiguous and y.flags.c_contiguous
        check(self.lib, self.lib.htb_add_vector_product(self.handle, trans.encode(), _ptr(a), _ptr(x), _ptr(b), _ptr(y), HTB_MEM_HOST))
        return y

    def add_matrix_product_row_major(self, trans, alpha, x, beta, y, mu):
        a, b = self._scalar(alpha), self._scalar(beta)
        assert x.dtype == self.dtype and y.dtype == self.dtype and x.flags.c_contiguous and y.flags.c_contiguous
        check(self.lib, self.lib.htb_add_matrix_product_row_major(self.handle, trans.encode(), _ptr(a), _ptr(x), _ptr(b), _ptr(y), mu, HTB_MEM_HOST))
        return y

    def set_permutations(self, target_perm, source_perm):
        tp = np.ascontiguousarray(target_perm, dtype=np.int32)
        sp = np.ascontiguousarray(source_perm, dtype=np.int32)
        check(self.lib, self.lib.htb_set_permutations(self.handle, _ptr(tp), _ptr(sp)))

    def add_vector_product_user_numbering(self, trans, alpha, x, beta, y):
        a, b = self._scalar(alpha), self._scalar(beta)
        check(self.lib, self.lib.htb_add_vector_product_user_numbering(self.handle, trans.encode(), _ptr(a), _ptr(x), _ptr(b), _ptr(y), HTB_MEM_HOST))
        return y

    def add_matrix_product_user_numbering(self, trans, alpha, x, beta, y, mu):
        a, b = self._scalar(alpha), self._scalar(beta)
        check(self.lib, self.lib.htb_add_matrix_product_user_numbering(self.handle, trans.encode(), _ptr(a), _ptr(x), _ptr(b), _ptr(y), mu, HTB_MEM_HOST))
        return y

    # raw device pointers (HTB_MEM_DEVICE), asynchronous on the handle's stream ---------------------
    def add_vector_product_device(self, trans, alpha, x_ptr: int, beta, y_ptr: int):
        a, b = self._scalar(alpha), self._scalar(beta)
        check(self.lib, self.lib.htb_add_vector_product(self.handle, trans.encode(), _ptr(a), C.c_void_p(x_ptr), _ptr(b), C.c_void_p(y_ptr), HTB_MEM_DEVICE))

    def add_matrix_product_row_major_device(self, trans, alpha, x_ptr: int, beta, y_ptr: int, mu: int):
        a, b = self._scalar(alpha), self._scalar(beta)
        check(self.lib, self.lib.htb_add_matrix_product_row_major(self.handle, trans.encode(), _ptr(a), C.c_void_p(x_ptr), _ptr(b), C.c_void_p(y_ptr), mu, HTB_MEM_DEVICE))

    # distributed -------------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes, world_size: int, rank: int, partition_offsets):
        po = np.ascontiguousarray(partition_offsets, dtype=np.int32)
        assert len(unique_id) == HTB_NCCL_UNIQUE_ID_BYTES and po.size == world_size + 1
        buf = C.create_string_buffer(unique_id, HTB_NCCL_UNIQUE_ID_BYTES)
        check(self.lib, self.lib.htb_comm_init(self.handle, C.cast(buf, C.c_void_p), world_size, rank, _ptr(po)))

    def dist_add_product_local_to_local(self, alpha, x, beta, y, mu=1, mem_kind=HTB_MEM_HOST, trans="N"):
        a, b = self._scalar(alpha), self._scalar(beta)
        xp = C.c_void_p(x) if isinstance(x, int) else _ptr(x)
        yp = C.c_void_p(y) if isinstance(y, int) else _ptr(y)
        check(self.lib, self.lib.htb_dist_add_product_local_to_local(self.handle, trans.encode(), _ptr(a), xp, _ptr(b), yp, mu, mem_kind))
        return y

    def dist_add_product_global_to_global(self, trans, alpha, x, beta, y, mu=1, mem_kind=HTB_MEM_HOST):
        a, b = self._scalar(alpha), self._scalar(beta)
        xp = C.c_void_p(x) if isinstance(x, int) else _ptr(x)
        yp = C.c_void_p(y) if isinstance(y, int) else _ptr(y)
        check(self.lib, self.lib.htb_dist_add_product_global_to_global(self.handle, trans.encode(), _ptr(a), xp, _ptr(b), yp, mu, mem_kind))
        return y


def _gmres(self, rhs, x, mem_kind=HTB_MEM_HOST, **options):
    """Device-resident restarted GMRES (htb_gmres). rhs / x: numpy arrays (host) or device addresses; x = initial guess in,
    solution out. options: restart, max_iterations, tolerance, orthogonalization, verbosity, compute_true_residual."""
    opt = htb_gmres_options()
    check(self.lib, self.lib.htb_gmres_default_options(C.byref(opt)))
    for k, v in options.items():
        assert hasattr(opt, k), k
        setattr(opt, k, v)
    res = htb_gmres_result()
    bp = C.c_void_p(rhs) if isinstance(rhs, int) else _ptr(rhs)
    xp = C.c_void_p(x) if isinstance(x, int) else _ptr(x)
    check(self.lib, self.lib.htb_gmres(self.handle, bp, xp, C.byref(opt), C.byref(res), mem_kind))
    return {k: getattr(res, k) for k, _ in res._fields_ if k != "reserved"}


Operator.gmres = _gmres


def nccl_unique_id() -> bytes:
    lib = load()
    buf = C.create_string_buffer(HTB_NCCL_UNIQUE_ID_BYTES)
    check(lib, lib.htb_nccl_get_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def host_register(array: np.ndarray):
    """Page-locks a numpy buffer for direct DMA by the host-pointer entry points (htb_host_register)."""
    lib = load()
    check(lib, lib.htb_host_register(C.c_void_p(array.ctypes.data), array.nbytes))


def host_unregister(array: np.ndarray):
    lib = load()
    check(lib, lib.htb_host_unregister(C.c_void_p(array.ctypes.data)))


def set_option(key: str, value: int):
    lib = load()
    check(lib, lib.htb_set_option(key.encode(), int(value)))


def get_option(key: str) -> int:
    lib = load()
    v = C.c_int64(0)
    check(lib, lib.htb_get_option(key.encode(), C.byref(v)))
    return int(v.value)


def leaves_from_arrays(rows, cols, m, n, rank, flags, data0, data1):
    """Builds a ctypes htb_leaf array from numpy columns (data0/data1 are integer addresses)."""
    k = len(rows)
    arr = (htb_leaf * k)()
    view = np.frombuffer(arr, dtype=LEAF_NP_DTYPE)
    view["row_offset"], view["col_offset"], view["nb_rows"], view["nb_cols"] = rows, cols, m, n
    view["rank"], view["flags"], view["data0"], view["data1"] = rank, flags, data0, data1
    return arr


LEAF_NP_DTYPE = np.dtype(
    [
        ("row_offset", np.int32),
        ("col_offset", np.int32),
        ("nb_rows", np.int32),
        ("nb_cols", np.int32),
        ("rank", np.int32),
        ("flags", np.int32),
        ("data0", np.uint64),
        ("data1", np.uint64),
    ]
)
assert LEAF_NP_DTYPE.itemsize == C.sizeof(htb_leaf)
