#!/usr/bin/env python
"""tools/tune.py — packer / launch option sweep on one GPU (development aid, not the bench).

Assembles the workload ONCE with the reference (oracle/_ref, as bench.py does), then for every option set
packs + uploads a fresh leaf store and times device-resident products with CUDA events on the launching
stream, printing one JSON line per option set (total ms, per-pass ms, achieved GB/s, parity vs the reference).

  python tools/tune.py --n 1000000 --set ring_stages=3 --set ring_stages=4,stage_bytes=32768 ...
"""
from __future__ import annotations

import argparse
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
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--dtype", default="double", choices=["double", "complex"])
    ap.add_argument("--symmetry", default="N", choices=["N", "S"])
    ap.add_argument("--trans", default="N")
    ap.add_argument("--mu", type=int, default=1)
    ap.add_argument("--partitions", type=int, default=1, help="tune ONE row strip of a P-way distributed operator on one GPU")
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--set", action="append", default=[], help="comma separated key=value list; one product configuration per --set")
    args = ap.parse_args()

    import torch

    import bench
    from htool_b200 import capi
    from oracle import refharness as R

    R.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    case = R.RefCase(**(bench.case_kwargs(args.n, args.dtype, args.symmetry, args.partitions, args.rank) if args.partitions > 1 else bench.case_kwargs(args.n, args.dtype, args.symmetry)))
    print(json.dumps({"assembly_s": time.perf_counter() - t0, **{k: v for k, v in case.info().items() if k in ("nb_leaves", "coefficients", "coefficients_twice")}}), flush=True)
    dtype = case.np_dtype
    esize = np.dtype(dtype).itemsize
    ni, no = (case.nb_cols, case.nb_rows) if args.trans == "N" else (case.nb_rows, case.nb_cols)
    mu = args.mu
    x = bench.seeded_x(ni * mu, dtype)
    y_ref = np.zeros(no * mu, dtype)
    if mu == 1:
        case.vector_product(args.trans, 1.0, x, 0.0, y_ref, variant="openmp")
    else:
        case.matrix_product_row_major(args.trans, 1.0, x, 0.0, y_ref, mu, variant="openmp")
    defaults = {k: capi.get_option(k) for k in ("block_rows", "piece_cols", "stage_bytes", "cseg_bytes", "ring_stages", "reduce_ring_stages", "evict_first", "m_ring_stages", "m_reduce_ring_stages", "mrhs_min", "fused_symmetric", "target_block_rows", "tail_split", "pdl", "reduce_blocks_per_cta", "sort_units", "m_b_ring_log2", "m_reduce_warps", "m_pad", "m_stage_input", "m_small_runs", "m_reduce_split", "m_b_producers", "m_near_field", "m_nf_rows", "m_b_global", "m_fast_tall")}
    stream = torch.cuda.Stream()
    x_d = torch.from_numpy(x).cuda()
    tdt = torch.float64 if dtype == np.float64 else torch.complex128
    for spec in args.set or [""]:
        opts = dict(defaults)
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            opts[k] = int(v)
        try:
            for k, v in opts.items():
                capi.set_option(k, v)
            t0 = time.perf_counter()
            case.desc.device = 0
            op = capi.Operator(case.desc)
            t_pack = time.perf_counter() - t0
            info = op.info()
            y = np.zeros(no * mu, dtype)

            def product(xp, yp, device):
                if mu == 1:
                    (op.add_vector_product_device if device else op.add_vector_product)(args.trans, 1.0, xp, 0.0, yp)
                else:
                    (op.add_matrix_product_row_major_device if device else op.add_matrix_product_row_major)(args.trans, 1.0, xp, 0.0, yp, mu)

            product(x, y, False)
            parity = float(np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref))
            op.set_stream(stream.cuda_stream)
            y_d = torch.zeros(no * mu, dtype=tdt, device="cuda")
            torch.cuda.synchronize()
            for _ in range(3):
                product(x_d.data_ptr(), y_d.data_ptr(), True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                product(x_d.data_ptr(), y_d.data_ptr(), True)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            op.profile_passes(True)
            for _ in range(5):
                product(x_d.data_ptr(), y_d.data_ptr(), True)
            pt = op.pass_times()
            op.profile_passes(False)
            nbytes = esize * (info["coefficients"] + mu * (ni + no))
            flops = 2.0 * mu * (info["coefficients"] + info["coefficients_twice"]) * (1 if dtype == np.float64 else 4)
            print(json.dumps({"opts": {k: v for k, v in opts.items() if v != defaults[k]}, "mu": mu, "ms": ms, "gbs": nbytes / ms / 1e6, "tflops": flops / ms / 1e9, "parity": parity,
                              "passes_ms": {k: v["ms"] / 5 for k, v in pt.items()}, "pack_s": t_pack, "store_gb": info["store_bytes"] / 1e9, "blocks": [info["nb_target_blocks"], info["nb_source_blocks"]],
                              "workspace_gb": info["workspace_bytes"] / 1e9, "descriptor_mb": info["descriptor_bytes"] / 1e6}), flush=True)
            op.set_stream(None)
            op.close()
        except Exception as ex:  # keep sweeping
            print(json.dumps({"opts": spec, "error": str(ex)}), flush=True)
        finally:
            for k, v in defaults.items():
                capi.set_option(k, v)


if __name__ == "__main__":
    main()
