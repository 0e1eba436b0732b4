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

    # ---- reference results of every distributed product (oracle/dist_parity.py: the reference's strip products + the
    # MPI collectives of its DistributedOperator linalg replaced by torch.distributed on CPU tensors) --------------------
    from oracle.dist_parity import DistReference, as_real, nccl_sweep

    cpu = dist.new_group(backend="gloo") if args.backend == "nccl" else None
    tt = torch.from_numpy
    ref = DistReference(case, world, rank, offsets, mu, sym=args.sym, cpu_group=cpu)
    rng = np.random.default_rng(6)
    x_global, x_local, alpha, beta, y0, y0_global = ref.x_global, ref.x_local, ref.alpha, ref.beta, ref.y0, ref.y0_global
    transposes, ref_l2l, ref_g2g, prod = ref.transposes, ref.ref_l2l, ref.ref_g2g, ref.prod
    lo_e, hi_e = ref.lo_e, ref.hi_e
    y_ref = ref_l2l["N"]

    def err(y, r=None):
        return ref.err(y, y_ref if r is None else r)

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
        sweep = nccl_sweep(op, ref, capi)
        if rank == 0:
            print("dist_worker sweep:", {k: f"{v:.1e}" for k, v in sweep.items()}, flush=True)
        errs.extend(sweep.values())
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
