"""oracle/dist_parity.py — TEST INFRASTRUCTURE. Parity sweep of the row-sharded (one process per GPU) products against
the reference, shared by tests/dist_worker.py and bench.py (N > 1: the sweep runs BEFORE any timing and a failure
aborts the run). Never imported by the product.

What the reference's DistributedOperator linalg does around the per-rank operator is reproduced with the reference's
own strip products (oracle/_ref, RestrictedGlobalToLocalHMatrix on the strip HMatrix) and torch.distributed collectives
on CPU tensors in place of MPI (MPI is absent from the image):
  l2l 'N'   (add_distributed_operator_vector_product_local_to_local.hpp:19-46): Allgatherv of x, strip product;
  l2l T/C   (:47-87): z_r = alpha op(H_r)^T x_r (global length), Alltoallv of the slices, out = beta out + sum_r z_r[own
            slice] in rank order;
  g2g 'N'   (add_distributed_operator_vector_product_global_to_global.hpp:43-76): own rows of out = beta out + alpha H_r x,
            Allgatherv;
  g2g T/C   (:51-57,77-83): Allreduce(sum) of z_r, + beta out;
and their row-major matrix twins (…matrix_product_row_major_local_to_local.hpp:25-95, …row_major_global_to_global.hpp:18-84).
"""
from __future__ import annotations

import numpy as np


def transposes_for(dtype, sym):
    if dtype == np.float64 or sym == "S":
        return ["T"]
    if sym == "N":
        return ["T", "C"]
    return ["C"]


def as_real(a):
    return a.view(np.float64) if a.dtype == np.complex128 else a


class DistReference:
    """Reference results of every distributed product for one (strip case, mu): computed once, compared many times."""

    def __init__(self, case, world, rank, offsets, mu, sym="N", cpu_group=None, seed=5):
        import torch
        import torch.distributed as dist

        self.case, self.world, self.rank, self.offsets, self.mu = case, world, rank, offsets, mu
        dtype = case.np_dtype
        self.dtype = dtype
        n_local, n_global = case.nb_rows, case.nb_cols
        self.n_local, self.n_global = n_local, n_global
        rng = np.random.default_rng(seed)  # same global vectors on every rank; each rank only USES its slice
        xg = rng.random(n_global * mu) - 0.5
        if dtype == np.complex128:
            xg = xg + 1j * (rng.random(n_global * mu) - 0.5)
        self.x_global = xg.astype(dtype)
        lo_e, hi_e = int(offsets[rank]) * mu, int(offsets[rank + 1]) * mu
        self.lo_e, self.hi_e = lo_e, hi_e
        self.x_local = np.ascontiguousarray(self.x_global[lo_e:hi_e])
        self.alpha, self.beta = (0.7, -1.3) if dtype == np.float64 else (0.7 + 0.2j, -1.3 + 0.4j)
        self.y0 = (rng.random(n_local * mu) - 0.5).astype(dtype)
        self.y0_global = (rng.random(n_global * mu) - 0.5).astype(dtype)
        self.transposes = transposes_for(dtype, sym)
        tt = torch.from_numpy
        alpha, beta = self.alpha, self.beta

        self.ref_l2l = {"N": self.prod("N", alpha, self.x_global, beta, self.y0.copy())}
        self.ref_g2g = {}
        for t in self.transposes:
            z = self.prod(t, alpha, self.x_local, 0.0, np.zeros(n_global * mu, dtype))
            send = [tt(as_real(np.ascontiguousarray(z[int(offsets[r]) * mu: int(offsets[r + 1]) * mu]))) for r in range(world)]
            recv = None
            # (gloo has no all_to_all: every slice owner gathers its slices instead)
            for r in range(world):
                got = [torch.zeros_like(send[r]) for _ in range(world)] if rank == r else None
                dist.gather(send[r], got, dst=r, group=cpu_group)
                if rank == r:
                    recv = got
            out = beta * self.y0
            for r in range(world):
                out = out + recv[r].numpy().view(dtype)
            self.ref_l2l[t] = out
            zsum = tt(as_real(z.copy()))
            dist.all_reduce(zsum, group=cpu_group)
            self.ref_g2g[t] = zsum.numpy().view(dtype) + beta * self.y0_global
        yl = self.prod("N", alpha, self.x_global, beta, self.y0_global[lo_e:hi_e].copy())
        w = 2 if dtype == np.complex128 else 1
        parts = [torch.zeros(int(offsets[r + 1] - offsets[r]) * mu * w, dtype=torch.float64) for r in range(world)]
        dist.all_gather(parts, tt(as_real(yl)), group=cpu_group)
        self.ref_g2g["N"] = torch.cat(parts).numpy().view(dtype)

    def prod(self, trans, a, x, b, y):
        if self.mu == 1:
            self.case.vector_product(trans, a, x, b, y, variant="global_to_local_operator")
        else:
            self.case.matrix_product_row_major(trans, a, x, b, y, self.mu, variant="global_to_local_operator")
        return y

    @staticmethod
    def err(y, ref):
        return float(np.linalg.norm(y - ref) / np.linalg.norm(ref))


def nccl_sweep(op, ref: DistReference, capi, repeats=5, host=True, device=True):
    """Every distributed entry point of the C ABI on `op` (htb_comm_init done) against `ref`. Returns {path: worst error}."""
    import torch

    mu, alpha, beta, y0, dtype = ref.mu, ref.alpha, ref.beta, ref.y0, ref.dtype
    n_local = ref.n_local
    errs = {}

    def note(key, e):
        errs[key] = max(errs.get(key, 0.0), e)

    for it in range(repeats):  # repeated: gather buffers (double-buffered by epoch), flags and events are reused across calls
        y = y0.copy()
        xs = ref.x_local * (1.0 + it)  # a different x every time: a stale buffer would show
        op.dist_add_product_local_to_local(alpha, xs, beta, y, mu)
        note("l2l_N_host", ref.err((y - beta * y0) / (1.0 + it) + beta * y0, ref.ref_l2l["N"]))
    if host:
        # page-locked host vectors (htb_host_register): zero copy when mu == 1 and the gather goes through peer memory
        x_pin, y_pin = ref.x_local.copy(), y0.copy()
        capi.host_register(x_pin)
        capi.host_register(y_pin)
        op.dist_add_product_local_to_local(alpha, x_pin, beta, y_pin, mu)
        note("l2l_N_pinned_host", ref.err(y_pin, ref.ref_l2l["N"]))
        capi.host_unregister(x_pin)
        capi.host_unregister(y_pin)
    x_d = torch.from_numpy(ref.x_local).cuda()
    if device:
        y_d = torch.from_numpy(y0.copy()).cuda()
        op.dist_add_product_local_to_local(alpha, x_d.data_ptr(), beta, y_d.data_ptr(), mu, capi.HTB_MEM_DEVICE)
        op.synchronize()
        note("l2l_N_device", ref.err(y_d.cpu().numpy(), ref.ref_l2l["N"]))
    for t in ref.transposes:  # T / C local-to-local: exchange of the slices + rank-ordered sum
        y = y0.copy()
        op.dist_add_product_local_to_local(alpha, ref.x_local, beta, y, mu, trans=t)
        note(f"l2l_{t}_host", ref.err(y, ref.ref_l2l[t]))
        if device:
            y_d = torch.from_numpy(y0.copy()).cuda()
            op.dist_add_product_local_to_local(alpha, x_d.data_ptr(), beta, y_d.data_ptr(), mu, capi.HTB_MEM_DEVICE, trans=t)
            op.synchronize()
            note(f"l2l_{t}_device", ref.err(y_d.cpu().numpy(), ref.ref_l2l[t]))
        y = np.full(n_local * mu, np.nan, dtype)  # beta == 0 ignores out
        op.dist_add_product_local_to_local(alpha, ref.x_local, 0.0, y, mu, trans=t)
        note(f"l2l_{t}_beta0", ref.err(y, ref.ref_l2l[t] - beta * y0))
    xg_d = torch.from_numpy(ref.x_global).cuda()
    for t in ["N"] + ref.transposes:  # global-to-global
        y = ref.y0_global.copy()
        op.dist_add_product_global_to_global(t, alpha, ref.x_global, beta, y, mu)
        note(f"g2g_{t}_host", ref.err(y, ref.ref_g2g[t]))
        if device:
            yg_d = torch.from_numpy(ref.y0_global.copy()).cuda()
            op.dist_add_product_global_to_global(t, alpha, xg_d.data_ptr(), beta, yg_d.data_ptr(), mu, capi.HTB_MEM_DEVICE)
            op.synchronize()
            note(f"g2g_{t}_device", ref.err(yg_d.cpu().numpy(), ref.ref_g2g[t]))
    return errs
