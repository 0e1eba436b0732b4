// htool_b200/csrc/mkernels.cu — multi-RHS (mu >= 8, double) sm_100a kernels: the leaves become batched dense
// contractions on the FP64 tensor cores (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4, the only FP64 MMA shape on
// sm_100a; there is no FP64 tcgen05).
//
// What they replace in the reference: openmp_internal_add_hmatrix_matrix_product_row_major
// (include/htool/hmatrix/linalg/add_hmatrix_matrix_product_row_major.hpp:112-178): one gemm per dense leaf
// (matrix/linalg/add_matrix_matrix_product_row_major.hpp:23-46) and two per low-rank leaf with a heap temp
// a(rank * mu) (hmatrix/lrmat/linalg/add_lrmat_matrix_product_row_major.hpp:11-27). B and C are ROW-major with mu
// contiguous, exactly as the reference's row-major kernels take them.
//
// Same store, same streams, same TMA ring as the single-RHS kernels (kernels.cu); the right-hand sides are handled in
// groups of MC <= 64 columns (VS = MC rounded up to 8 is the vector stride of the scratch):
//   REDUCE_M  T[k][c] = sum_i P[i][k] X[i][c]   one unit per warp. The block's X rows sit in shared memory (row stride
//             VS + 8 doubles: conflict-free A fragments); the warp walks the 8 column tiles c with one P fragment, so
//             the bank conflicts of the unpadded column-major panel are paid once per 8 DMMAs. 32 accumulators / lane.
//   APPLY_M   C[i][c] += sum_k P[i][k] T[k][c]  the 8 warps split the ROWS of the block (16 rows each) and keep their
//             16 x 64 slice of C in registers for the whole block (deterministic: one owner per C entry, fixed unit
//             order). A unit is touched only by the warps whose rows it meets; its panel fragment is reused for the
//             8 column tiles, the T / X fragments come straight from global memory (L2).
//   COMBINE_M sums the per-chunk partials of pieces with several producer chunks into TF.
// Roofline: FP64 tensor pipe (measured 37.2 TFLOP/s DMMA on B200, profiles/r01_fp64_peak_b200.json);
// flops = 2 * mu * C per product.
#include "mkernels.cuh"

#include <cstdint>

namespace htb {

namespace {

constexpr int kApplyWarps    = 16; // APPLY_M: the consumer warps split the rows of the block (8 rows each), 1 CTA / SM
constexpr int kReduceWarps   = 16; // REDUCE_M: one unit per warp, 1 CTA / SM (the X block takes 76 KiB of shared memory)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// ---- mbarrier / bulk copy (same protocol as kernels.cu) --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity), "r"(0x989680u)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

struct MSide {
    const BlockDesc *blocks;
    const StageDesc *stages;
    const uint32_t *order;
    const unsigned char *stream;
    const MUnit *munits;
    int block_rows, stage_bytes, ring_stages;
};

struct RingPos {
    uint32_t slot, phase;
    __device__ __forceinline__ RingPos() : slot(0), phase(0) {}
    __device__ __forceinline__ void advance(uint32_t ring) {
        if (++slot == ring) {
            slot = 0;
            phase ^= 1u;
        }
    }
};

__device__ __forceinline__ void produce(const MSide &ks, const BlockDesc &bd, unsigned char *ring, uint64_t *full, uint64_t *empty, int twice_only) {
    const uint64_t policy = l2_evict_first_policy();
    if (bd.n_stages == 0)
        return;
    RingPos pos;
    StageDesc next = ks.stages[bd.first_stage];
    for (uint32_t st = 0; st < bd.n_stages; st++) {
        const StageDesc sd = next;
        if (st + 1 < bd.n_stages)
            next = ks.stages[bd.first_stage + st + 1];
        if (twice_only && !(sd.flags & 1u))
            continue;
        mbar_wait(smem_u32(&empty[pos.slot]), pos.phase ^ 1u);
        mbar_arrive_expect_tx(smem_u32(&full[pos.slot]), sd.nbytes);
        bulk_g2s(smem_u32(ring + static_cast<size_t>(pos.slot) * ks.stage_bytes), ks.stream + sd.byte_off, sd.nbytes, smem_u32(&full[pos.slot]), policy);
        pos.advance(ks.ring_stages);
    }
}

__device__ __forceinline__ void init_barriers(int ring, uint64_t *full, uint64_t *empty, int consumer_warps) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < ring; s++) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), consumer_warps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

// ---- REDUCE_M -------------------------------------------------------------------------------------------------------
// smem: [ring | Xs (block_rows x (VS + 8)) | barriers]
__global__ void __launch_bounds__((kReduceWarps + 1) * 32) reduce_m_kernel(MSide ks, MArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const BlockDesc bd         = ks.blocks[ks.order[blockIdx.x]];
    const uint32_t n_my_stages = a.twice_only ? bd.n_twice_stages : bd.n_stages;
    if (n_my_stages == 0)
        return;
    const int XS        = a.vs + 8;
    unsigned char *ring = smem_raw;
    double *Xs          = reinterpret_cast<double *>(smem_raw + static_cast<size_t>(ks.ring_stages) * ks.stage_bytes);
    uint64_t *full      = reinterpret_cast<uint64_t *>(Xs + static_cast<size_t>(ks.block_rows + 4) * XS);
    uint64_t *empty     = full + ks.ring_stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;

    init_barriers(ks.ring_stages, full, empty, kReduceWarps);
    // the block's rows of the input matrix, columns [col0, col0 + mc), zero padded to VS columns; 4 zero rows follow
    // the block (the last k-step of a unit that ends the block reads up to 3 rows past it)
    for (int idx = threadIdx.x; idx < (ks.block_rows + 4) * a.vs; idx += (kReduceWarps + 1) * 32) {
        const int i = idx / a.vs, c = idx - i * a.vs;
        const long long gr = static_cast<long long>(bd.row_start) + i + a.in_shift;
        Xs[i * XS + c]     = (i < bd.nrows && c < a.mc && gr >= 0 && gr < a.in_rows) ? a.in[gr * a.ld_in + a.col0 + c] : 0.;
    }
    __syncthreads();

    if (warp == kReduceWarps) {
        if (lane == 0)
            produce(ks, bd, ring, full, empty, a.twice_only);
        return;
    }

    const int MT = a.vs >> 3; // column tiles
    RingPos pos;
    uint32_t ubase = warp;
    for (uint32_t q = 0; q < n_my_stages; q++, pos.advance(ks.ring_stages)) {
        mbar_wait(smem_u32(&full[pos.slot]), pos.phase);
        const unsigned char *stage = ring + static_cast<size_t>(pos.slot) * ks.stage_bytes;
        const StageHeader hdr      = *reinterpret_cast<const StageHeader *>(stage);
        const Unit *units          = reinterpret_cast<const Unit *>(stage + sizeof(StageHeader));
        const double *data         = reinterpret_cast<const double *>(stage + hdr.data_byte_off);
        const MUnit *mun           = ks.munits + hdr.first_unit;
        // (ADDVEC units need nothing here: direction 0 reads the x rows of dense leaves from the input directly)
        uint32_t out_next = ubase < hdr.n_panel ? mun[ubase].out : 0u;
        for (uint32_t u = ubase; u < hdr.n_panel; u += kReduceWarps) {
            const Unit un      = units[u];
            const uint32_t out = out_next;
            if (u + kReduceWarps < hdr.n_panel)
                out_next = mun[u + kReduceWarps].out; // in flight during this unit's contractions
            if (a.twice_only && !unit_twice(un.geom))
                continue;
            const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom), w = unit_w(un.geom);
            const uint32_t ld = unit_ld(h, sizeof(double));
            const double *P   = data + un.data_off;
            const int NT      = (w + 7) >> 3;
            double acc[8][2][2];
#pragma unroll
            for (int mt = 0; mt < 8; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
                    acc[mt][nt][0] = acc[mt][nt][1] = 0.;
            const double *xrow = Xs + (row0 + tig) * XS + g;
            for (uint32_t i0 = 0; i0 < h; i0 += 4) {
                // B fragments: P[i0 + tig][8 nt + g], zero outside the unit
                const uint32_t i  = i0 + tig;
                const bool iv     = i < h;
                const uint32_t ic = iv ? i : h - 1;
                double b[2];
                b[0] = (iv && static_cast<uint32_t>(g) < w) ? P[(static_cast<uint32_t>(g) < w ? g : w - 1) * ld + ic] : 0.;
                b[1] = 0.;
                if (NT > 1)
                    b[1] = (iv && 8u + g < w) ? P[(8u + g < w ? 8u + g : w - 1) * ld + ic] : 0.;
                const double *xr = xrow + i0 * XS;
#pragma unroll
                for (int mt = 0; mt < 8; mt++) {
                    if (mt < MT) {
                        const double av = xr[8 * mt]; // X[row0 + i0 + tig][8 mt + g]; rows past the unit meet b == 0
                        dmma(acc[mt][0], av, b[0]);
                        if (NT > 1)
                            dmma(acc[mt][1], av, b[1]);
                    }
                }
            }
            // D[c = 8 mt + g][k = 8 nt + 2 tig + j] -> T[k][c]
            double *T = a.mscratch + static_cast<size_t>(out) * a.vsp;
#pragma unroll
            for (int mt = 0; mt < 8; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const uint32_t k = 8u * nt + 2u * tig + j;
                        if (mt < MT && k < w)
                            T[static_cast<size_t>(k) * a.vsp + 8 * mt + g] = acc[mt][nt][j];
                    }
        }
        ubase = (ubase - hdr.n_panel) & (kReduceWarps - 1);
        __syncwarp();
        if (lane == 0)
            mbar_arrive(smem_u32(&empty[pos.slot]));
    }
}

// ---- APPLY_M --------------------------------------------------------------------------------------------------------
// The 16 consumer warps split the ROWS of the block: warp w owns the 8-row tile w
// for all the columns of the group and keeps that slice of C in registers for the whole block: one owner per C
// entry, units in stream order -> deterministic. A unit is only touched by the warps whose rows it meets (a dense
// 8 x 8 leaf: one warp, a 128 x 8 panel: all of them, each on its own rows), so no coefficient is re-read.
//   D[i][c] += A[i][k] B[k][c]:  A = panel fragment P[8 t + g][4 s + tig], read straight from the ring slot and
//   reused for the 8 column tiles (its 4-way bank conflict is paid once per 8 DMMAs); B = T[4 s + tig][8 ct + g].
// The T pieces of the low-rank units are needed by every warp that meets the unit: the PRODUCER WARP bulk-copies them
// from TF (row stride VS + 8 doubles: conflict-free B fragments) into the c area of the ring slot, next to the stage,
// on the same mbarrier; the offsets go to a small table in the slot. Dense leaves read their rows of the input matrix
// from global memory (nobody else needs them); so do the low-rank units that did not fit the c area.
// smem: [ring: slot = stage | c area | offset table] [barriers]
constexpr uint32_t kCAreaBytes  = 16384;
constexpr uint32_t kCoffEntries = 512; // units of a stage that can have a staged T piece
constexpr uint32_t kSlotExtra   = kCAreaBytes + kCoffEntries * 2;

__device__ __forceinline__ void produce_apply_m(const MSide &ks, const BlockDesc &bd, unsigned char *ring, uint32_t slot_bytes, uint64_t *full, uint64_t *empty, const MArgs &a, int lane) {
    const uint64_t policy = l2_evict_first_policy();
    if (bd.n_stages == 0)
        return;
    RingPos pos;
    for (uint32_t st = 0; st < bd.n_stages; st++) {
        const StageDesc sd = ks.stages[bd.first_stage + st]; // same address in every lane: one transaction
        if (a.twice_only && !(sd.flags & 1u))
            continue;
        unsigned char *slot = ring + static_cast<size_t>(pos.slot) * slot_bytes;
        uint16_t *coff      = reinterpret_cast<uint16_t *>(slot + ks.stage_bytes + kCAreaBytes);
        const uint32_t bar  = smem_u32(&full[pos.slot]);
        if (lane == 0)
            mbar_wait(smem_u32(&empty[pos.slot]), pos.phase ^ 1u);
        __syncwarp();
        // pass 1: where does every low-rank unit's T piece go (prefix sum over the units, 32 at a time)?
        const MUnit *mun = ks.munits + sd.first_unit;
        uint32_t used = 0, total = 0;
        const uint32_t n_panel = sd.n_panel < kCoffEntries ? sd.n_panel : kCoffEntries;
        for (uint32_t u0 = 0; u0 < n_panel; u0 += 32) {
            const uint32_t u = u0 + lane;
            uint32_t bytes   = 0, src = 0;
            if (u < n_panel) {
                const MUnit mu = mun[u];
                if (mu.flags & 2u) {
                    bytes = ((mu.flags >> 8) & 0xffu) * a.vsp * 8u;
                    src   = mu.src;
                }
            }
            // inclusive scan
            uint32_t incl = bytes;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d)
                    incl += v;
            }
            const uint32_t off = used + incl - bytes;
            const bool fits    = bytes > 0 && off + bytes <= kCAreaBytes;
            // a piece that does not fit is read from global memory by the consumers; later (smaller) pieces may still fit,
            // but keep it simple: everything after the first overflow of this chunk is also left out
            const uint32_t ok_mask = __ballot_sync(0xffffffffu, bytes == 0 || fits);
            const bool staged      = fits && (ok_mask == 0xffffffffu || lane < __ffs(static_cast<int>(~ok_mask)) - 1);
            if (u < n_panel)
                coff[u] = staged ? static_cast<uint16_t>(off >> 3) : static_cast<uint16_t>(0xffffu);
            const uint32_t staged_bytes = staged ? bytes : 0u;
            uint32_t sum = staged_bytes;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1)
                sum += __shfl_xor_sync(0xffffffffu, sum, d);
            total += sum;
            // remember what to copy: issue after the expect_tx below (second pass recomputes, cheap)
            used += __shfl_sync(0xffffffffu, incl, 31);
            (void)src;
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive_expect_tx(bar, sd.nbytes + total);
        __syncwarp();
        if (lane == 0)
            bulk_g2s(smem_u32(slot), ks.stream + sd.byte_off, sd.nbytes, bar, policy);
        // pass 2: the copies (one per lane)
        for (uint32_t u0 = 0; u0 < n_panel; u0 += 32) {
            const uint32_t u = u0 + lane;
            if (u < n_panel) {
                const uint16_t o = coff[u];
                if (o != 0xffffu) {
                    const MUnit mu = mun[u];
                    const uint32_t bytes = ((mu.flags >> 8) & 0xffu) * a.vsp * 8u;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(slot + ks.stage_bytes + static_cast<uint32_t>(o) * 8u)),
                                 "l"(a.mscratch + static_cast<size_t>(mu.src) * a.vsp), "r"(bytes), "r"(bar)
                                 : "memory");
                }
            }
        }
        __syncwarp();
        pos.advance(ks.ring_stages);
    }
}

template <int RT, bool FULL>
__global__ void __launch_bounds__((kApplyWarps + 1) * 32) apply_m_kernel(MSide ks, MArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const BlockDesc bd = ks.blocks[ks.order[blockIdx.x]];
    if (a.twice_only && !(bd.flags & 1u))
        return;
    const uint32_t n_my_stages = a.twice_only ? bd.n_twice_stages : bd.n_stages;
    const uint32_t slot_bytes  = ks.stage_bytes + kSlotExtra;
    unsigned char *ring = smem_raw;
    uint64_t *full      = reinterpret_cast<uint64_t *>(smem_raw + static_cast<size_t>(ks.ring_stages) * slot_bytes);
    uint64_t *empty     = full + ks.ring_stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;

    init_barriers(ks.ring_stages, full, empty, kApplyWarps);
    __syncthreads();

    if (warp == kApplyWarps) {
        produce_apply_m(ks, bd, ring, slot_bytes, full, empty, a, lane);
        return;
    }

    const int CT        = FULL ? 8 : a.vs >> 3;                // column tiles of the group (FULL: mc == 64)
    const uint32_t rlo  = 8u * RT * warp, rhi = rlo + 8u * RT; // my rows of the block
    double acc[RT][8][2];                                      // C[rlo + 8 t + g][8 ct + 2 tig + j]
#pragma unroll
    for (int t = 0; t < RT; t++)
#pragma unroll
        for (int ct = 0; ct < 8; ct++)
            acc[t][ct][0] = acc[t][ct][1] = 0.;

    RingPos pos;
    for (uint32_t q = 0; q < n_my_stages; q++, pos.advance(ks.ring_stages)) {
        mbar_wait(smem_u32(&full[pos.slot]), pos.phase);
        const unsigned char *stage = ring + static_cast<size_t>(pos.slot) * slot_bytes;
        const StageHeader hdr      = *reinterpret_cast<const StageHeader *>(stage);
        const Unit *units          = reinterpret_cast<const Unit *>(stage + sizeof(StageHeader));
        const double *data         = reinterpret_cast<const double *>(stage + hdr.data_byte_off);
        const double *carea        = reinterpret_cast<const double *>(stage + ks.stage_bytes);
        const uint16_t *coff       = reinterpret_cast<const uint16_t *>(stage + ks.stage_bytes + kCAreaBytes);
        const MUnit *mun           = ks.munits + hdr.first_unit;
        if (rlo < static_cast<uint32_t>(bd.nrows)) {
            for (uint32_t u = 0; u < hdr.n_units; u++) {
                const Unit un = units[u];
                const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom);
                if (row0 >= rhi || row0 + h <= rlo || (a.twice_only && !unit_twice(un.geom)))
                    continue; // not my rows
                if (u >= hdr.n_panel) {
                    // ADDVEC (direction 1): C rows += z, the TF vectors produced by REDUCE_M over side 0
                    const uint32_t src = mun[u].src;
#pragma unroll
                    for (int t = 0; t < RT; t++) {
                        const uint32_t row = rlo + 8u * t + g;
                        if (row >= row0 && row < row0 + h) {
                            const double *z = a.mscratch + static_cast<size_t>(src + row - row0) * a.vsp + 2 * tig;
#pragma unroll
                            for (int ct = 0; ct < 8; ct++)
                                if (FULL || ct < CT) {
                                    const double2 v = *reinterpret_cast<const double2 *>(z + 8 * ct);
                                    acc[t][ct][0] += v.x;
                                    acc[t][ct][1] += v.y;
                                }
                        }
                    }
                    continue;
                }
                const uint32_t w  = unit_w(un.geom);
                const uint32_t ld = unit_ld(h, sizeof(double));
                const double *P   = data + un.data_off;
                // B fragments: row (4 s + tig) of the unit's input vectors, column 8 ct + g
                const uint32_t co = u < kCoffEntries ? coff[u] : 0xffffu;
                const double *Bsrc;
                long long brows; // rows available from Bsrc[0]
                size_t bld;
                if (co != 0xffffu) { // staged T piece (shared memory)
                    Bsrc  = carea + co;
                    bld   = a.vsp;
                    brows = w;
                } else {
                    const uint32_t src = mun[u].src;
                    if (src & 0x80000000u) { // dense leaf: rows of the input matrix
                        const long long r0 = static_cast<long long>(src & 0x7fffffffu) + a.in_shift;
                        Bsrc  = a.in + r0 * a.ld_in + a.col0;
                        bld   = a.ld_in;
                        brows = a.in_rows - r0;
                    } else {
                        Bsrc  = a.mscratch + static_cast<size_t>(src) * a.vsp;
                        bld   = a.vsp;
                        brows = w;
                    }
                }
                for (uint32_t k0 = 0; k0 < w; k0 += 4) {
                    const uint32_t k = k0 + tig;
                    const bool kv    = k < w && static_cast<long long>(k) < brows;
                    const double *Bk = Bsrc + static_cast<size_t>(kv ? k : 0u) * bld + g;
                    double bf[8];
#pragma unroll
                    for (int ct = 0; ct < 8; ct++)
                        bf[ct] = (kv && (FULL || (ct < CT && 8 * ct + g < a.mc))) ? Bk[8 * ct] : 0.;
                    const double *Pk = P + (k < w ? k : w - 1) * ld;
#pragma unroll
                    for (int t = 0; t < RT; t++) {
                        const uint32_t row = rlo + 8u * t + g;
                        const bool rv      = row >= row0 && row < row0 + h;
                        const double af    = (rv && k < w) ? Pk[rv ? row - row0 : 0u] : 0.;
                        if (rlo + 8u * t < row0 + h && rlo + 8u * t + 8u > row0) { // tile meets the unit (warp-uniform)
#pragma unroll
                            for (int ct = 0; ct < 8; ct++)
                                if (FULL || ct < CT)
                                    dmma(acc[t][ct], af, bf[ct]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(smem_u32(&empty[pos.slot]));
    }
    // epilogue: alpha / beta, one write per C entry
#pragma unroll
    for (int t = 0; t < RT; t++) {
        const int i = static_cast<int>(rlo) + 8 * t + g;
        if (i < bd.nrows) {
            const long long gr = static_cast<long long>(bd.row_start) + i + a.out_shift;
            if (gr >= 0 && gr < a.out_rows) {
                double *o = a.out + gr * a.ld_out + a.col0;
#pragma unroll
                for (int ct = 0; ct < 8; ct++)
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const int c = 8 * ct + 2 * tig + j;
                        if (FULL || (ct < CT && c < a.mc)) {
                            double r = a.alpha * acc[t][ct][j];
                            if (!a.beta_is_zero)
                                r = fma(a.beta, o[c], r);
                            o[c] = r;
                        }
                    }
            }
        }
    }
}

// One warp per (piece, vector of the piece): sums the per-chunk partial vectors in chunk order (fixed summation order).
__global__ void combine_m_kernel(const CombineEntry *entries, int n, double *mscratch, int vs, int vsp, int twice_only) {
    const int warps_per_block = blockDim.x >> 5;
    const long long gw        = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
    const int e               = static_cast<int>(gw >> 5);
    const uint32_t k          = static_cast<uint32_t>(gw & 31);
    if (e >= n)
        return;
    const CombineEntry ce = entries[e];
    const uint32_t lane = threadIdx.x & 31, len = combine_len(ce.packed), n_sum = combine_n_sum(ce.packed);
    if (k >= len || (twice_only && !combine_twice(ce.packed)))
        return;
    for (uint32_t c = lane; c < static_cast<uint32_t>(vs); c += 32) {
        const double *p = mscratch + (static_cast<size_t>(ce.src) + k) * vsp + c;
        double v        = 0.;
#pragma unroll 4
        for (uint32_t j = 0; j < n_sum; j++)
            v += p[static_cast<size_t>(j) * len * vsp];
        mscratch[(static_cast<size_t>(ce.dst_first) + k) * vsp + c] = v;
    }
}

inline MSide make_mside(const SideDevice &s, const LaunchConfig &cfg, int ring) {
    return MSide{s.blocks, s.stages, s.order, s.stream, s.munits, cfg.block_rows, cfg.stage_bytes, ring};
}

} // namespace

size_t reduce_m_smem_bytes(const LaunchConfig &cfg, int vs) {
    return static_cast<size_t>(cfg.m_reduce_ring_stages) * cfg.stage_bytes + sizeof(double) * (cfg.block_rows + 4) * (vs + 8) + 16 * static_cast<size_t>(cfg.m_reduce_ring_stages);
}
size_t apply_m_smem_bytes(const LaunchConfig &cfg) {
    return static_cast<size_t>(cfg.m_ring_stages) * (cfg.stage_bytes + kSlotExtra) + 16 * static_cast<size_t>(cfg.m_ring_stages);
}

cudaError_t configure_mkernels(const LaunchConfig &cfg) {
    auto set = [](const void *f, size_t smem) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        // several CTAs per SM: ask for the largest shared-memory carve-out, the default only guarantees one block
        return e != cudaSuccess ? e : cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    };
    cudaError_t e;
    if ((e = set(reinterpret_cast<const void *>(reduce_m_kernel), reduce_m_smem_bytes(cfg, 64))) != cudaSuccess)
        return e;
    const void *apply[2] = {reinterpret_cast<const void *>(apply_m_kernel<1, false>), reinterpret_cast<const void *>(apply_m_kernel<1, true>)};
    for (const void *f : apply)
        if ((e = set(f, apply_m_smem_bytes(cfg))) != cudaSuccess)
            return e;
    return cudaSuccess;
}

cudaError_t launch_reduce_m(const SideDevice &side, const LaunchConfig &cfg, const MArgs &args, cudaStream_t stream) {
    if (side.n_blocks == 0)
        return cudaSuccess;
    reduce_m_kernel<<<side.n_blocks, (kReduceWarps + 1) * 32, reduce_m_smem_bytes(cfg, args.vs), stream>>>(make_mside(side, cfg, cfg.m_reduce_ring_stages), args);
    return cudaGetLastError();
}

cudaError_t launch_apply_m(const SideDevice &side, const LaunchConfig &cfg, const MArgs &args, cudaStream_t stream) {
    if (side.n_blocks == 0)
        return cudaSuccess;
    MArgs a        = args;
    a.beta_is_zero = args.beta == 0. ? 1 : 0;
    const MSide ms     = make_mside(side, cfg, cfg.m_ring_stages);
    const size_t smem  = apply_m_smem_bytes(cfg);
    const int threads  = (kApplyWarps + 1) * 32;
    const bool full    = args.mc == 64;
    // 16 consumer warps x one 8-row tile each = up to 128 block rows (warps past the block's rows only help nobody)
    if (full)
        apply_m_kernel<1, true><<<side.n_blocks, threads, smem, stream>>>(ms, a);
    else
        apply_m_kernel<1, false><<<side.n_blocks, threads, smem, stream>>>(ms, a);
    return cudaGetLastError();
}

cudaError_t launch_combine_m(const SideDevice &side, double *mscratch, int vs, int twice_only, cudaStream_t stream) {
    if (side.n_combine_m == 0)
        return cudaSuccess;
    const int warps = 8; // 32 warps (one per vector of the piece, pieces hold <= 32 vectors) per entry
    combine_m_kernel<<<static_cast<unsigned>((static_cast<long long>(side.n_combine_m) * 32 + warps - 1) / warps), warps * 32, 0, stream>>>(side.combine_m, side.n_combine_m, mscratch, vs, vs + 8, twice_only);
    return cudaGetLastError();
}

} // namespace htb
