#!/usr/bin/env python
"""tools/assembly_time.py — device leaf assembly (htb_create_compressed) timed on one GPU (development aid, not the bench).

Assembles the workload once with the reference (for the block cluster tree and the comparison), then creates the operator from
the stripped descriptor once per --set option list and prints the timing breakdown of htb_get_compression_info.
  HTB_PACK_TIMING=1 python tools/assembly_time.py --n 1000000 --set upload_headers_only=1 --set upload_headers_only=0
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--set", action="append", default=[])
    args = ap.parse_args()
    import bench
    from htool_b200 import capi
    from oracle import refharness as R

    R.set_num_threads(os.cpu_count() or 1)
    case = R.RefCase(**bench.case_kwargs(args.n, "double", "N"))
    info = case.info()
    print(json.dumps({"reference_assembly_s": info["build_seconds"], "threads": info["omp_threads"], "nb_leaves": info["nb_leaves"]}), flush=True)
    lv = case.leaves().copy()
    ref_rank = lv["rank"].copy()
    lv["rank"] = np.where(ref_rank >= 0, capi.HTB_RANK_COMPRESS, -1)
    lv["data0"], lv["data1"] = 0, 0
    arr = (capi.htb_leaf * max(1, len(lv))).from_buffer_copy(lv.tobytes())
    d = capi.htb_hmatrix_desc()
    C.memmove(C.byref(d), C.byref(case.desc), C.sizeof(capi.htb_hmatrix_desc))
    d.leaves = C.cast(arr, C.POINTER(capi.htb_leaf))
    d.device = 0
    tp, sp = case.points(0), case.points(1)
    x = bench.seeded_x(case.nb_cols, np.float64)
    y_ref = np.zeros(case.nb_rows)
    op0 = capi.Operator(case.desc)
    op0.add_vector_product("N", 1.0, x, 0.0, y_ref)
    op0.close()
    for spec in args.set or [""]:
        opts = {k: int(v) for k, v in (kv.split("=") for kv in spec.split(",") if kv)}
        saved = {k: capi.get_option(k) for k in opts}
        for k, v in opts.items():
            capi.set_option(k, v)
        try:
            t0 = time.perf_counter()
            op = capi.Operator(d, generator=("laplace_reg", tp, sp, 0.0), compress_epsilon=1e-4)
            t = time.perf_counter() - t0
            ci = op.compression_info()
            y = np.zeros(case.nb_rows)
            op.add_vector_product("N", 1.0, x, 0.0, y)
            same = bool(np.array_equal(op.leaf_ranks(), ref_rank))
            op.close()
        finally:
            for k, v in saved.items():
                capi.set_option(k, v)
        print(json.dumps({"opts": opts, "create_s": t, "bit_identical_product": bool(np.array_equal(y, y_ref)), "same_ranks": same,
                          **{k: v for k, v in ci.items() if k.startswith("seconds")}}), flush=True)


if __name__ == "__main__":
    main()
