#!/usr/bin/env python
"""tools/pack_time.py — the packer's layout pass timed on the host, no GPU (development aid, not the bench).

Assembles the workload once with the reference (for the block cluster tree and the ranks; the leaf list is cached under
/tmp), strips the coefficient pointers as the device assembly does (htb_create_compressed packs a descriptor whose leaves
carry ranks and no data: only the layout and the stage headers are computed), calls htb_pack_host for both sides and
prints seconds and a digest of every table, so that a packer change can be checked to be byte-identical:
  HTB_PACK_TIMING=1 python tools/pack_time.py --n 1000000 --reps 3
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def leaf_list(n, dtype, symmetry):
    import bench
    from oracle import refharness as R

    cache = f"/tmp/htb_leaves_{dtype}_{symmetry}_{n}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        return z["leaves"], z["desc"].tobytes()
    R.set_num_threads(os.cpu_count() or 1)
    case = R.RefCase(**bench.case_kwargs(n, dtype, symmetry))
    lv = case.leaves().copy()
    d = case.desc
    np.savez(cache, leaves=lv, desc=np.frombuffer(bytes(d), dtype=np.uint8))
    return leaf_list(n, dtype, symmetry)


def digest(ptr, nbytes):
    if not ptr or nbytes <= 0:
        return "-"
    return hashlib.sha1(C.string_at(ptr, nbytes)).hexdigest()[:12]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dtype", default="double", choices=["double", "complex"])
    ap.add_argument("--symmetry", default="N", choices=["N", "S"])
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--set", action="append", default=[], help="packer option key=value")
    ap.add_argument("--with-data", action="store_true", help="keep the reference's coefficients (no cache): times the fill pass of the host-packed path too")
    ap.add_argument("--lib", default=None, help="another build of libhtool_b200.so (A/B of a packer change)")
    args = ap.parse_args()
    from htool_b200 import capi

    if args.lib:
        capi.LIB_PATH = os.path.abspath(args.lib)
    lib = capi.load()
    case = None
    if args.with_data:
        import bench
        from oracle import refharness as R

        R.set_num_threads(os.cpu_count() or 1)
        case = R.RefCase(**bench.case_kwargs(args.n, args.dtype, args.symmetry))  # (kept alive: the leaves point into it)
        lv, desc_bytes = case.leaves().copy(), bytes(case.desc)
    else:
        lv, desc_bytes = leaf_list(args.n, args.dtype, args.symmetry)
        lv = lv.copy()
        lv["data0"], lv["data1"] = 0, 0
    for kv in args.set:
        k, v = kv.split("=")
        capi.set_option(k, int(v))
    capi.set_option("pack_generate_dense", 1)  # leaves without data are legal: their coefficients are produced on the device
    arr = (capi.htb_leaf * max(1, len(lv))).from_buffer_copy(lv.tobytes())
    d = capi.htb_hmatrix_desc.from_buffer_copy(desc_bytes)
    d.leaves = C.cast(arr, C.POINTER(capi.htb_leaf))
    for rep in range(args.reps):
        out = {"rep": rep, "leaves": len(lv)}
        for side in (0, 1):
            p = capi.htb_packed_side()
            t0 = time.perf_counter()
            capi.check(lib, lib.htb_pack_host(C.byref(d), side, C.byref(p)))
            out[f"seconds_side{side}"] = round(time.perf_counter() - t0, 3)
            out[f"digest_side{side}"] = {
                "blocks": digest(p.blocks, 32 * p.n_blocks), "stages": digest(p.stages, 32 * p.n_stages), "order": digest(p.order, 4 * p.n_blocks),
                "combine": digest(p.combine, 16 * p.n_combine), "combine_dst": digest(p.combine_dst, 8 * p.n_combine_dst),
                "munits": digest(p.munits, 16 * p.n_munits), "combine_m": digest(p.combine_m, 16 * p.n_combine_m),
                "aux_reduce": digest(p.aux_reduce, p.aux_bytes), "aux_apply": digest(p.aux_apply, p.aux_bytes),
                "dense_tasks": digest(p.dense_tasks, 32 * p.n_dense_tasks), "lowrank_tasks": digest(p.lowrank_tasks, 32 * p.n_lowrank_tasks),
                "headers": digest(p.headers, p.header_bytes), "header_offsets": digest(p.header_offsets, 8 * (p.n_stages + 1)),
                "scalars": [p.n_blocks, p.n_stages, p.stream_bytes, p.scratch_elems, p.mscratch_elems, p.aux_bytes, p.header_bytes],
            }
            capi.check(lib, lib.htb_pack_free(C.byref(p)))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
