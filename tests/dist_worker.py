"""tests/dist_worker.py — one rank of the row-sharded product (launched by torch.distributed.run from the tests).

Every rank lets the reference assemble ITS row strip (build(gen, target, source, rank, rank),
distributed_operator/utility.hpp:56), and computes y_local = H_strip * allgather(x_local):

  --backend gloo (CPU, no GPU needed): the gather goes through torch.distributed (gloo) and the strip product through
      tests/stream_emulator.py on the bytes htb_create would upload — this checks the HOST side of the distributed
      path (partition offsets, strip descriptors, local/remote split of the source blocks) without a device;
  --backend nccl (one GPU per rank): htb_comm_init + htb_dist_add_product_local_to_local, i.e. the product's own NCCL
      gather overlapped with the local-source leaves, with host and with device pointers.

Both compare with the reference's product on the same strip object and (rank 0) with the undistributed reference
operator. Exit code 0 = all ranks within tolerance.
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--points", type=int, default=3000)
    ap.add_argument("--scalar", default="double")
    ap.add_argument("--sym", default="N")
    ap.add_argument("--rhs", type=int, default=1)
    ap.add_argument("--p2p", type=int, default=1, help="0: NCCL gather of x instead of the peer-memory push (htb_set_option dist_p2p)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from oracle import refharness as R

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if args.backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend="gloo")
    R.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    kw = dict(n=args.points, n_partitions=world, partition_rank=rank)
    if args.scalar == "complex":
        kw.update(dtype="complex", kernel="helmholtz")
    if args.sym != "N":
        kw.update(symmetry=args.sym, uplo="L")
    case = R.RefCase(**kw)
    dtype, mu = case.np_dtype, args.rhs
    n_local, n_global = case.nb_rows, case.nb_cols
    dev = "cuda" if args.backend == "nccl" else "cpu"

    # partition offsets from the strips themselves (PartitionFromCluster::get_offset_of_partition)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = n_local
    dist.all_reduce(sizes)
    offsets = np.concatenate([[0], np.cumsum(sizes.cpu().numpy())]).astype(np.int32)
    assert offsets[-1] == n_global and case.desc.row_offset - case.desc.col_offset == offsets[rank]

    rng = np.random.default_rng(5)  # same global x on every rank; each rank only USES its slice
    x_global = rng.random(n_global * mu) - 0.5
    if dtype == np.complex128:
        x_global = x_global + 1j * (rng.random(n_global * mu) - 0.5)
    x_global = x_global.astype(dtype)
    x_local = np.ascontiguousarray(x_global[offsets[rank] * mu: offsets[rank + 1] * mu])
    alpha, beta = (0.7, -1.3) if dtype == np.float64 else (0.7 + 0.2j, -1.3 + 0.4j)
    y0 = (rng.random(n_local * mu) - 0.5).astype(dtype)
    y_ref = y0.copy()
    if mu == 1:
        case.vector_product("N", alpha, x_global, beta, y_ref, variant="global_to_local_operator")
    else:
        case.matrix_product_row_major("N", alpha, x_global, beta, y_ref, mu, variant="global_to_local_operator")

    def err(y, ref=None):
        ref = y_ref if ref is None else ref
        return float(np.linalg.norm(y - ref) / np.linalg.norm(ref))

    # ---- reference results of the transposed local-to-local and the global-to-global products -----------------------
    # What the reference's DistributedOperator linalg does around the per-rank operator, with the MPI collectives
    # replaced by torch.distributed on CPU tensors (gloo):
    #   l2l T/C (add_distributed_operator_vector_product_local_to_local.hpp:47-87): z_r = alpha op(H_r)^T x_r (global
    #     length), Alltoallv of the slices, out = beta out + sum_r z_r[own slice] in rank order;
    #   g2g N (…global_to_global.hpp:43-76): own rows of out = beta out + alpha H_r x, Allgatherv;
    #   g2g T/C (:51-83): Allreduce(sum) of z_r, + beta out.
    cpu = dist.new_group(backend="gloo") if args.backend == "nccl" else None
    tt = torch.from_numpy

    def prod(trans, a, x, b, y):
        if mu == 1:
            case.vector_product(trans, a, x, b, y, variant="global_to_local_operator")
        else:
            case.matrix_product_row_major(trans, a, x, b, y, mu, variant="global_to_local_operator")
        return y

    def as_real(a):
        return a.view(np.float64) if a.dtype == np.complex128 else a

    transposes = ["T"] if (dtype == np.float64 or args.sym == "S") else ["C"]
    if dtype == np.complex128 and args.sym == "N":
        transposes = ["T", "C"]
    y0_global = (rng.random(n_global * mu) - 0.5).astype(dtype)  # same on every rank
    lo_e, hi_e = int(offsets[rank]) * mu, int(offsets[rank + 1]) * mu
    ref_l2l, ref_g2g = {}, {}
    for t in transposes:
        z = prod(t, alpha, x_local, 0.0, np.zeros(n_global * mu, dtype))
        send = [tt(as_real(np.ascontiguousarray(z[int(offsets[r]) * mu: int(offsets[r + 1]) * mu]))) for r in range(world)]
        recv = [torch.zeros(as_real(x_local).size, dtype=torch.float64) for _ in range(world)]
        # gloo has no all_to_all: every slice owner gathers its slices instead
        for r in range(world):
            got = [torch.zeros_like(send[r]) for _ in range(world)] if rank == r else None
            dist.gather(send[r], got, dst=r, group=cpu)
            if rank == r:
                recv = got
        out = beta * y0
        for r in range(world):
            out = out + recv[r].numpy().view(dtype)
        ref_l2l[t] = out
        zsum = tt(as_real(z.copy()))
        dist.all_reduce(zsum, group=cpu)
        ref_g2g[t] = zsum.numpy().view(dtype) + beta * y0_global
    yl = prod("N", alpha, x_global, beta, y0_global[lo_e:hi_e].copy())
    parts = [torch.zeros(int(offsets[r + 1] - offsets[r]) * mu * (2 if dtype == np.complex128 else 1), dtype=torch.float64) for r in range(world)]
    dist.all_gather(parts, tt(as_real(yl)), group=cpu)
    ref_g2g["N"] = torch.cat(parts).numpy().view(dtype)

    errs = []
    if args.backend == "gloo":
        from stream_emulator import Emulator
        from oracle.flatcase import FlatCase

        # the gather of x (MPI_Allgatherv in the reference, linalg/utility.hpp:27), unequal counts
        pieces = [torch.zeros(int(offsets[r + 1] - offsets[r]) * mu, dtype=torch.from_numpy(x_local).dtype) for r in range(world)]
        dist.all_gather(pieces, torch.from_numpy(x_local))
        gathered = torch.cat(pieces).numpy()
        assert np.array_equal(gathered, x_global)
        em = Emulator(FlatCase.from_desc(case.desc))
        assert mu == 1
        y = y0.copy()
        assert em.vector_product("N", alpha, gathered, beta, y) == 0
        errs.append(err(y))
        # local / remote split of the source blocks (dist.cu: htb_comm_init): every source block is either inside the
        # rank's own partition or not; the local ones only need x_local
        side1 = em.side[1]
        lo, hi = int(offsets[rank]), int(offsets[rank + 1])
        local = [(int(b["row_start"]) >= lo and int(b["row_start"] + b["nrows"]) <= hi) for b in side1.blocks]
        # (a block that straddles the partition boundary counts as remote: it waits for the gather)
        n_local_blocks = sum(local)
        assert 0 < n_local_blocks < len(local) or world == 1
        # the local blocks read nothing outside x_local: zero the remote part of x, their REDUCE output must not change
        x_masked = gathered.copy()
        x_masked[:lo] = 0
        x_masked[hi:] = 0
        s_full, s_mask = em.new_scratch(), em.new_scratch()
        order = side1.order
        side1.order = np.array([b for b in order if local[b]], dtype=order.dtype)
        em.reduce(1, gathered, 0, s_full, False, False)
        em.reduce(1, x_masked, 0, s_mask, False, False)
        side1.order = order
        assert np.array_equal(np.nan_to_num(s_full), np.nan_to_num(s_mask))
        # transposed local-to-local and global-to-global: strip product through the emulator, exchange through gloo
        for t in transposes:
            z = np.zeros(n_global, dtype)
            assert em.vector_product(t, alpha, x_local, 0.0, z) == 0
            out = beta * y0
            for r in range(world):
                piece = tt(as_real(np.ascontiguousarray(z[offsets[r]: offsets[r + 1]])))
                got = [torch.zeros_like(piece) for _ in range(world)] if rank == r else None
                dist.gather(piece, got, dst=r)
                if rank == r:
                    for g in got:
                        out = out + g.numpy().view(dtype)
            errs.append(err(out, ref_l2l[t]))
            zs = tt(as_real(z.copy()))
            dist.all_reduce(zs)
            errs.append(err(zs.numpy().view(dtype) + beta * y0_global, ref_g2g[t]))
        yl = y0_global[lo_e:hi_e].copy()
        assert em.vector_product("N", alpha, gathered, beta, yl) == 0
        parts = [torch.zeros(int(offsets[r + 1] - offsets[r]) * (2 if dtype == np.complex128 else 1), dtype=torch.float64) for r in range(world)]
        dist.all_gather(parts, tt(as_real(yl)))
        errs.append(err(torch.cat(parts).numpy().view(dtype), ref_g2g["N"]))
    else:
        from htool_b200 import capi

        case.desc.device = local_rank
        capi.set_option("dist_p2p", args.p2p)
        op = capi.Operator(case.desc)
        uid = torch.zeros(capi.HTB_NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        op.comm_init(bytes(uid.cpu().numpy().tobytes()), world, rank, offsets)
        assert op.info()["dist_gather"] == (2 if args.p2p else 1), op.info()["dist_gather"]  # NVLink box: peer mappings must work
        for it in range(5):  # repeated: gather buffers (double-buffered by epoch), flags and events are reused across calls
            y = y0.copy()
            xs = x_local * (1.0 + it)  # a different x every time: a stale buffer would show
            op.dist_add_product_local_to_local(alpha, xs, beta, y, mu)
            errs.append(err((y - beta * y0) / (1.0 + it) + beta * y0))
        # page-locked host vectors (htb_host_register): zero copy when mu == 1 and the gather goes through peer memory
        x_pin, y_pin = x_local.copy(), y0.copy()
        capi.host_register(x_pin)
        capi.host_register(y_pin)
        op.dist_add_product_local_to_local(alpha, x_pin, beta, y_pin, mu)
        errs.append(err(y_pin))
        capi.host_unregister(x_pin)
        capi.host_unregister(y_pin)
        x_d, y_d = torch.from_numpy(x_local).cuda(), torch.from_numpy(y0.copy()).cuda()
        op.dist_add_product_local_to_local(alpha, x_d.data_ptr(), beta, y_d.data_ptr(), mu, capi.HTB_MEM_DEVICE)
        op.synchronize()
        errs.append(err(y_d.cpu().numpy()))
        tdt = torch.float64 if dtype == np.float64 else torch.complex128
        for t in transposes:  # T / C local-to-local: grouped send/recv of the slices + rank-ordered sum
            y = y0.copy()
            op.dist_add_product_local_to_local(alpha, x_local, beta, y, mu, trans=t)
            errs.append(err(y, ref_l2l[t]))
            y_d = torch.from_numpy(y0.copy()).cuda()
            op.dist_add_product_local_to_local(alpha, x_d.data_ptr(), beta, y_d.data_ptr(), mu, capi.HTB_MEM_DEVICE, trans=t)
            op.synchronize()
            errs.append(err(y_d.cpu().numpy(), ref_l2l[t]))
            y = np.full(n_local * mu, np.nan, dtype)  # beta == 0 ignores out
            op.dist_add_product_local_to_local(alpha, x_local, 0.0, y, mu, trans=t)
            errs.append(err(y, ref_l2l[t] - beta * y0))
        xg_d = torch.from_numpy(x_global).cuda()
        for t in ["N"] + transposes:  # global-to-global
            y = y0_global.copy()
            op.dist_add_product_global_to_global(t, alpha, x_global, beta, y, mu)
            errs.append(err(y, ref_g2g[t]))
            yg_d = torch.from_numpy(y0_global.copy()).cuda()
            op.dist_add_product_global_to_global(t, alpha, xg_d.data_ptr(), beta, yg_d.data_ptr(), mu, capi.HTB_MEM_DEVICE)
            op.synchronize()
            errs.append(err(yg_d.cpu().numpy(), ref_g2g[t]))
        if mu == 1:
            # distributed device-resident GMRES (htb_gmres): local slices in / out, inner products summed over the ranks.
            # Checker: the numpy oracle on the GLOBAL operator, its matvec = the reference's strip products gathered over gloo.
            from oracle.gmres_oracle import gmres

            def mv(v):
                yl = prod("N", 1.0, np.ascontiguousarray(v.astype(dtype)), 0.0, np.zeros(n_local, dtype))
                parts = [torch.zeros(int(offsets[r + 1] - offsets[r]) * (2 if dtype == np.complex128 else 1), dtype=torch.float64) for r in range(world)]
                dist.all_gather(parts, tt(as_real(yl)), group=cpu)
                return torch.cat(parts).numpy().view(dtype)

            b_global = (rng.random(n_global) - 0.5).astype(dtype)
            for iters, restart in [(5, 40), (7, 3)]:
                xo, io = gmres(mv, b_global, restart=restart, max_iterations=iters, tolerance=0.0, reorthogonalize=True)
                xl = np.zeros(n_local, dtype)
                ig = op.gmres(np.ascontiguousarray(b_global[offsets[rank]: offsets[rank + 1]]), xl, restart=restart, max_iterations=iters, tolerance=0.0,
                              orthogonalization=capi.HTB_GMRES_CGS2)
                assert ig["iterations"] == io["iterations"] == iters and ig["matvecs"] == io["matvecs"], (ig, io)
                assert abs(ig["true_relative_residual"] - io["true_relative_residual"]) <= 1e-9 * max(1.0, io["true_relative_residual"]), (ig, io)
                errs.append(1e-4 * err(xl, xo[offsets[rank]: offsets[rank + 1]]))  # iterates agree to 1e-8
        op.close()

    worst = torch.tensor([max(errs)], dtype=torch.float64, device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dist_worker backend={args.backend} world={world} n={args.points} dtype={args.scalar} sym={args.sym} mu={mu}: worst rel. l2 error {worst.item():.3e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if worst.item() < 1e-12 else 1)


if __name__ == "__main__":
    main()
