#!/usr/bin/env python
"""tools/gmres_time.py — times htb_gmres solves on the bench workload (development aid): per-solve wall time, matvecs,
and the same number of bare products for comparison."""
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
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--set", action="append", default=[""])
    args = ap.parse_args()
    import torch

    import bench
    from htool_b200 import capi
    from oracle import refharness as R

    R.set_num_threads(os.cpu_count() or 1)
    case = R.RefCase(**bench.case_kwargs(argparse.Namespace(n=args.n, dtype="double", symmetry="N", mu=1, gpus=1)))
    case.desc.device = 0
    op = capi.Operator(case.desc)
    n = case.nb_rows
    stream = torch.cuda.Stream()
    op.set_stream(stream.cuda_stream)
    b = torch.from_numpy(bench.seeded_x(n, np.float64, seed=2)).cuda()
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for spec in args.set:
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            capi.set_option(k, int(v))
        op.gmres(b.data_ptr(), x.data_ptr(), mem_kind=capi.HTB_MEM_DEVICE, max_iterations=3, tolerance=0.0)
        times, info = [], None
        for _ in range(args.reps):
            x.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info = op.gmres(b.data_ptr(), x.data_ptr(), mem_kind=capi.HTB_MEM_DEVICE, tolerance=1e-10)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(info["matvecs"]):
            op.add_vector_product_device("N", 1.0, b.data_ptr(), 0.0, y.data_ptr())
        torch.cuda.synchronize()
        bare = time.perf_counter() - t0
        print(json.dumps({"opts": spec, "solve_ms": [round(1e3 * t, 3) for t in times], "matvecs": info["matvecs"], "iterations": info["iterations"],
                          "bare_products_ms": round(1e3 * bare, 3), "fraction": bare / min(times)}), flush=True)
    op.close()


if __name__ == "__main__":
    main()
