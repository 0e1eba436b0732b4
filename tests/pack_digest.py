"""tests/pack_digest.py — prints one digest of everything htb_pack_host produces for a seeded synthetic leaf list (both sides:
stream bytes and every table). tests/test_packer.py runs it under several OMP_NUM_THREADS: the layout must not depend on how
many threads the packer runs on (its serial passes run one side per thread, its block loops are dynamic)."""
import ctypes as C
import hashlib
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    from htool_b200 import capi
    from oracle.flatcase import random_flatcase

    dtype_code, symmetric = int(sys.argv[1]), (sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None)
    flat = random_flatcase(seed=11, dtype_code=dtype_code, nb_rows=2600, nb_cols=2300, n_leaves=1400, max_dim=200, max_rank=20, symmetric=symmetric)
    lib = capi.load()
    h = hashlib.sha1()
    for side in (0, 1):
        p = capi.htb_packed_side()
        capi.check(lib, lib.htb_pack_host(C.byref(flat.desc), side, C.byref(p)))
        for ptr, nbytes in [(p.blocks, 32 * p.n_blocks), (p.stages, 32 * p.n_stages), (p.order, 4 * p.n_blocks), (p.combine, 16 * p.n_combine), (p.combine_dst, 8 * p.n_combine_dst),
                            (p.stream, p.stream_bytes), (p.munits, 16 * p.n_munits), (p.combine_m, 16 * p.n_combine_m), (p.aux_reduce, p.aux_bytes), (p.aux_apply, p.aux_bytes)]:
            h.update(str(nbytes).encode())
            if ptr and nbytes > 0:
                h.update(C.string_at(ptr, nbytes))
        h.update(repr((p.n, p.n_blocks, p.n_stages, p.scratch_elems, p.cs_base, p.cs_elems, p.part_base, p.part_elems, p.mscratch_elems)).encode())
        capi.check(lib, lib.htb_pack_free(C.byref(p)))
    print("digest", h.hexdigest())


if __name__ == "__main__":
    main()
