// htool_b200/csrc/mkernels.cu — multi-RHS (mu >= 2) sm_100a kernels: the leaves become batched dense contractions on
// the FP64 tensor cores (mma.sync m8n8k4 f64 = SASS DMMA.8x8x4, the only FP64 MMA shape on sm_100a; there is no FP64
// tcgen05), for double AND complex<double> (2 x 2 real embedding, below).
//
// What they replace in the reference: openmp_internal_add_hmatrix_matrix_product_row_major
// (include/htool/hmatrix/linalg/add_hmatrix_matrix_product_row_major.hpp:112-178): one gemm per dense leaf
// (matrix/linalg/add_matrix_matrix_product_row_major.hpp:23-46, complex 'C' / hemm variants :49-84,114-139) and two per
// low-rank leaf with a heap temp a(rank * mu) (hmatrix/lrmat/linalg/add_lrmat_matrix_product_row_major.hpp:11-27). B and
// C are ROW-major with mu contiguous, exactly as the reference's row-major kernels take them.
//
// Same store, same streams, same TMA ring as the single-RHS kernels (kernels.cu). The right-hand sides are handled in
// groups of <= 64 REAL columns (32 complex ones); VS = that count rounded up to 8, VSP = VS + 8 the vector stride of the
// scratch. The unit of work is a RUN (store.hpp): the packer orders the units of a block by the rows they act on, so all
// the panels of one cluster are consecutive in the stream and form one column-major panel h x K with K in the hundreds;
// the kernels contract whole runs, no tile is padded along K and nothing is done per leaf. What every column of a run
// multiplies / produces comes from the stage's aux record (column table), bulk-copied next to the stage.
//   REDUCE_M  T[k][c] = sum_i P[i][k] X[i][c]. The block's rows of X sit in shared memory (row stride VSP: conflict-free A
//             fragments). A job = 8 columns of a run; the jobs are dealt round-robin over 24 consumer warps. 16
//             accumulators per lane. D[c][k] goes to the scratch vector the column table names.
//   APPLY_M   C[i][c] += sum_k P[i][k] B[k][c]. 16 consumer warps = 8 column tiles x 2 row-tile parities: EVERY warp works
//             on every run (its 8 columns, the 8-row tiles of its parity that the run meets), so a run of a small cluster
//             still occupies all the warps. C stays in registers for the whole block (one owner per entry, stream order:
//             deterministic). B rows (T vectors of low-rank columns, rows of the input matrix for dense columns) are
//             loaded straight from global memory one k-step ahead.
//   COMBINE_M sums the per-chunk partials of pieces with several producer chunks into TF.
// complex<double>: a complex panel h x K (re / im interleaved, ld = h) is the REAL panel 2h x K with ld = 2h. REDUCE_M
// contracts it against the real 2h x 2mu matrix whose row 2i is (re, im) of X[i][.] and row 2i+1 is (-im, re)
// (conjugated panel: (im, -re)): the result rows are the complex T[k][.] interleaved. APPLY_M doubles the contraction
// instead: A[i][2k + j] = (re, im)[j] of P[i][k] against B[2k][.] = T[k][.] and B[2k+1][.] = i T[k][.] (re / im swapped
// with a sign), so the accumulators hold complex C interleaved, exactly the row-major layout of the caller.
// Roofline: FP64 tensor pipe (measured 37.2 TFLOP/s DMMA on B200, profiles/r01_fp64_peak_b200.json);
// flops = 2 * mu * C per product, x 4 for complex.
#include "mkernels.cuh"

#include <algorithm>
#include <cstdint>

namespace htb {

namespace {

constexpr int kApplyWarps  = 16; // APPLY_M: 8 column tiles x 2 row-tile parities
constexpr int kReduceWarpsMax = 24; // REDUCE_M: jobs (8 columns of a run) dealt round-robin over the consumer warps (option m_reduce_warps)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
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
__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct MSide {
    const BlockDesc *blocks;
    const StageDesc *stages;
    const uint32_t *order;
    const unsigned char *stream;
    const unsigned char *aux; // aux_reduce or aux_apply of the side, by kernel
    const MUnit *munits;
    int block_rows, stage_bytes, aux_bytes, ring_stages;
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

// One lane streams the block's stages and their aux records into the ring (slot = [stage | aux record]).
__device__ __forceinline__ void produce(const MSide &ks, const BlockDesc &bd, unsigned char *ring, uint32_t slot_bytes, uint64_t *full, uint64_t *empty, int twice_only) {
    const uint64_t policy = l2_evict_first_policy();
    if (bd.n_stages == 0)
        return;
    RingPos pos;
    StageDesc next = ks.stages[bd.first_stage];
    for (uint32_t st = 0; st < bd.n_stages; st++) {
        const StageDesc sd = next;
        if (st + 1 < bd.n_stages)
            next = ks.stages[bd.first_stage + st + 1]; // in flight while this stage waits for its slot
        if (twice_only && !(sd.flags & 1u))
            continue;
        const uint32_t aux_len = static_cast<uint32_t>(sd.flags >> 1) * 16u;
        const uint32_t bar     = smem_u32(&full[pos.slot]);
        const uint32_t dst     = smem_u32(ring + static_cast<size_t>(pos.slot) * slot_bytes);
        mbar_wait(smem_u32(&empty[pos.slot]), pos.phase ^ 1u);
        mbar_arrive_expect_tx(bar, sd.nbytes + aux_len);
        bulk_g2s(dst, ks.stream + sd.byte_off, sd.nbytes, bar, policy);
        if (aux_len)
            bulk_g2s_plain(dst + ks.stage_bytes, ks.aux + static_cast<size_t>(sd.aux_off16) * 16u, aux_len, bar);
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

// One REDUCE_M job: acc[mt] += X-fragment(mt) x P-fragment over the k-steps (4 rows each) of a run, column tiles
// mt < mn of the right-hand sides (FULL: all 8, no predicates). b = P[i][col], zero outside the run: rows past the run meet
// b == 0. Plain loop on purpose: two resident warps per scheduler already keep the DMMA pipe full with the loads in front
// of each DMMA, and a software-pipelined variant (fragments of the next k-step in their own registers) measured SLOWER
// once the pipe is shared (tools/dmma_probe.cu: 27 vs 35 TFLOP/s at 8+ warps per SM: the register moves cost issue slots).
template <bool FULL>
__device__ __forceinline__ void reduce_job(double (&acc)[8][2], const double *Pc, const double *xj, int XS, uint32_t h, int tig, bool cv, int mn) {
    for (uint32_t i0 = 0; i0 < h; i0 += 4) {
        const uint32_t i = i0 + tig;
        const double b   = (cv && i < h) ? Pc[i < h ? i : h - 1u] : 0.;
        const double *xr = xj + i0 * XS;
#pragma unroll
        for (int mt = 0; mt < 8; mt++)
            if (FULL || mt < mn)
                dmma(acc[mt], xr[8 * mt], b); // X[row0 + i0 + tig][8 (m0 + mt) + g]
    }
}

// ---- REDUCE_M -------------------------------------------------------------------------------------------------------
// smem: [ring: slot = stage | aux] [Xs ((real rows of a block + 4) x VSP)] [barriers]
template <bool CPLX>
__global__ void __launch_bounds__((kReduceWarpsMax + 1) * 32) reduce_m_kernel(MSide ks, MArgs a) {
    constexpr int CS = CPLX ? 1 : 0; // a complex row is two real rows
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const BlockDesc bd         = ks.blocks[ks.order[blockIdx.x]];
    const uint32_t n_my_stages = a.twice_only ? bd.n_twice_stages : bd.n_stages;
    if (n_my_stages == 0)
        return;
    const uint32_t slot_bytes = ks.stage_bytes + ks.aux_bytes;
    const int XS              = a.vsp; // row stride of the X block in shared memory = 4 (mod 8): conflict-free A fragments (64-bit
                                       // loads are served per half warp: tig * XS + g, g < 4, must hit 16 different bank pairs)
    const int RB              = ks.block_rows << CS; // real rows of a block
    unsigned char *ring       = smem_raw;
    double *Xs                = reinterpret_cast<double *>(smem_raw + static_cast<size_t>(ks.ring_stages) * slot_bytes);
    uint64_t *full            = reinterpret_cast<uint64_t *>(Xs + static_cast<size_t>(RB + 4) * XS);
    uint64_t *empty           = full + ks.ring_stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const int kReduceWarps = static_cast<int>(blockDim.x >> 5) - 1; // consumer warps; the last warp is the producer

    uint64_t *xbar = empty + ks.ring_stages; // completion of the bulk copy of the X block
    // The staged group (capi.cu: rows one X-block row apart, 16 B aligned, zero padded) is ONE contiguous piece of memory
    // with the layout of the X block: the producer lane brings the block's rows with a single bulk copy that overlaps the
    // first stages, instead of ~11 dependent global loads per consumer thread in front of the first DMMA.
    const bool x_bulk       = !CPLX && a.ld_in == XS && a.col0_in == 0 && ((reinterpret_cast<uintptr_t>(a.in) & 15u) == 0);
    const long long gr0     = static_cast<long long>(bd.row_start) + a.in_shift; // input row of the block's first row
    const int x_lo          = gr0 < 0 ? static_cast<int>(gr0 < -bd.nrows ? bd.nrows : -gr0) : 0;
    const long long x_avail = a.in_rows - gr0;
    const int x_hi          = x_avail < x_lo ? x_lo : (x_avail < bd.nrows ? static_cast<int>(x_avail) : bd.nrows); // rows [x_lo, x_hi) of the block exist in the input

    init_barriers(ks.ring_stages, full, empty, kReduceWarps);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(xbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == kReduceWarps) {
        if (lane == 0) {
            if (x_bulk && x_hi > x_lo) {
                const uint32_t bytes = static_cast<uint32_t>(x_hi - x_lo) * static_cast<uint32_t>(XS) * 8u;
                mbar_arrive_expect_tx(smem_u32(xbar), bytes);
                bulk_g2s_plain(smem_u32(Xs + static_cast<size_t>(x_lo) * XS), a.in + (gr0 + x_lo) * a.ld_in, bytes, smem_u32(xbar));
            } else
                mbar_arrive(smem_u32(xbar));
            produce(ks, bd, ring, slot_bytes, full, empty, a.twice_only);
        }
        return;
    }
    // (the producer is already streaming) the block's rows of the input matrix, real columns [col0, col0 + mc), zero padded
    // to VS columns; 4 zero rows follow the block (the last k-step of a run that ends the block reads up to 3 rows past it)
    if (x_bulk) {
        const int nz = (RB + 4) - (x_hi - x_lo); // rows that the bulk copy does not bring: zeros
        for (int idx = threadIdx.x; idx < nz * a.vs; idx += kReduceWarps * 32) {
            int r       = idx / a.vs;
            const int c = idx - r * a.vs;
            r           = r < x_lo ? r : r + (x_hi - x_lo);
            Xs[r * XS + c] = 0.;
        }
    } else {
        for (int idx = threadIdx.x; idx < (RB + 4) * a.vs; idx += kReduceWarps * 32) {
            const int r = idx / a.vs, c = idx - r * a.vs;
            const int i = r >> CS;
            const long long gr = static_cast<long long>(bd.row_start) + i + a.in_shift;
            double v           = 0.;
            if (i < bd.nrows && c < a.mc && gr >= 0 && gr < a.in_rows) {
                const double *row = a.in + gr * a.ld_in + a.col0_in;
                if (!CPLX || !(r & 1))
                    v = row[c];
                else { // row 2i+1 of the embedding: i * X[i][.] = (-im, re); conjugated panel: -i * X[i][.] = (im, -re)
                    const double o = row[c ^ 1];
                    v              = ((c & 1) != 0) == (a.conj == 0) ? o : -o;
                }
            }
            Xs[r * XS + c] = v;
        }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(kReduceWarps * 32) : "memory");
    mbar_wait(smem_u32(xbar), 0);

    const int MT = a.vs >> 3; // column tiles of the right-hand sides
    RingPos pos;
    uint32_t jmod = 0; // jobs dealt so far, mod kReduceWarps
    for (uint32_t q = 0; q < n_my_stages; q++, pos.advance(ks.ring_stages)) {
        mbar_wait(smem_u32(&full[pos.slot]), pos.phase);
        const unsigned char *stage = ring + static_cast<size_t>(pos.slot) * slot_bytes;
        const StageHeader hdr      = *reinterpret_cast<const StageHeader *>(stage);
        const double *data         = reinterpret_cast<const double *>(stage + hdr.data_byte_off);
        const AuxHeader ah         = *reinterpret_cast<const AuxHeader *>(stage + ks.stage_bytes);
        const RunDesc *runs        = reinterpret_cast<const RunDesc *>(stage + ks.stage_bytes + sizeof(AuxHeader));
        const uint32_t *cols       = reinterpret_cast<const uint32_t *>(runs + ah.n_runs);
        for (uint32_t r = 0; r < ah.n_runs; r++) {
            const RunDesc rd = runs[r];
            if (a.twice_only && !(rd.flags & 1u))
                continue;
            const uint32_t K = rd.K, ntiles = (K + 7u) >> 3;
            const uint32_t hc = static_cast<uint32_t>(rd.h_minus_1) + 1u;
            const uint32_t h  = hc << CS;                                      // real rows
            const uint32_t ld = CPLX ? 2u * hc : unit_ld(hc, sizeof(double)); // real leading dimension
            const double *P   = data + (static_cast<size_t>(rd.data_off) << CS);
            const double *xrow = Xs + ((static_cast<uint32_t>(rd.row0) << CS) + tig) * XS + g;
            // jobs of this run: (tile of 8 columns) x (one of S groups of column tiles of the right-hand sides). S > 1 for tall
            // runs: a stage of a 122-row panel holds ~25 columns = 3 tiles, so with one job per tile a 4-stage ring feeds 12 of
            // the 24 warps; splitting the right-hand sides keeps every warp busy (the panel fragment is re-read from shared
            // memory by the S jobs, the DMMAs are the same). My jobs: jb = jb0, jb0 + W, ...; jb = S t + part.
            const uint32_t S     = (h >= 64u && a.reduce_split > 1 && MT >= a.reduce_split) ? static_cast<uint32_t>(a.reduce_split) : 1u;
            const uint32_t njobs = ntiles * S;
            const int mper       = (MT + static_cast<int>(S) - 1) / static_cast<int>(S); // column tiles per job
            uint32_t jb0 = static_cast<uint32_t>(warp + kReduceWarps) - jmod;
            if (jb0 >= static_cast<uint32_t>(kReduceWarps))
                jb0 -= kReduceWarps;
            jmod = (jmod + njobs) % static_cast<uint32_t>(kReduceWarps);
            for (uint32_t jb = jb0; jb < njobs; jb += kReduceWarps) {
                const uint32_t t = jb / S;
                const int m0     = static_cast<int>(jb - t * S) * mper;
                const int mn     = MT - m0 < mper ? MT - m0 : mper; // column tiles m0 .. m0 + mn - 1
                double acc[8][2];
#pragma unroll
                for (int mt = 0; mt < 8; mt++)
                    acc[mt][0] = acc[mt][1] = 0.;
                const uint32_t col  = 8u * t + g;
                const bool cv       = col < K;
                const double *Pc    = P + static_cast<size_t>(cv ? col : K - 1u) * ld;
                const double *xj    = xrow + 8 * m0;
                if (mn == 8)
                    reduce_job<true>(acc, Pc, xj, XS, h, tig, cv, 8);
                else
                    reduce_job<false>(acc, Pc, xj, XS, h, tig, cv, mn);
                // D[c = 8 (m0 + mt) + g][k = 8 t + 2 tig + j] -> scratch vector named by the column table
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const uint32_t k = 8u * t + 2u * tig + j;
                    if (k < K) {
                        double *T = a.mscratch + static_cast<size_t>(cols[rd.col0 + k]) * a.vsp + g + 8 * m0;
#pragma unroll
                        for (int mt = 0; mt < 8; mt++)
                            if (mt < mn)
                                T[8 * mt] = acc[mt][j];
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(smem_u32(&empty[pos.slot]));
    }
}

// ---- APPLY_M --------------------------------------------------------------------------------------------------------
// 16 consumer warps = CT column tiles x PAR = 16 / CT row-tile classes (CT = 8 for a full group of 64 real columns, fewer for
// small mu so that no warp multiplies zero columns). Warp w: column tile ct = w % CT (real columns 8 ct .. 8 ct + 7), class
// par = w / CT: it owns the 8-row tiles j PAR + par of the block for those columns,
// acc[j] = C[8 (j PAR + par) + g][8 ct + 2 tig + {0, 1}].
//   D[i][c] += A[i][k] B[k][c]:  A = panel fragment P[row][k0 + tig] read from the ring slot, one per row tile the run
//   meets; B = row (k0 + tig) of what the run's columns multiply, columns 8 ct + g.
// The B rows (T vectors REDUCE_M / COMBINE_M just wrote, rows of the input matrix for dense columns) form a STREAM in the
// order of the block's columns. Warp 17, the B producer, walks the column tables of the stages as they arrive and copies
// the rows with bulk copies (TMA; consecutive rows in ONE copy, see produce_b) into a ring of chunks of 32 rows whose row
// stride is the scratch's vector stride, completion counted in bytes on the chunk's mbarrier — the producer never waits
// for data and stays 2 - 3 chunks ahead of the consumers, so the DRAM latency of the B rows never meets a DMMA. Every run starts at a multiple of 4 in the stream and every stage at
// a multiple of 32: a k-step never straddles two chunks, a chunk never two stages.
// smem: [stage ring: slot = stage | aux] [B ring: chunk = 32 x VSP doubles] [barriers]
constexpr int kBChunk = 32;
constexpr int kBProducersMax = 3; // B producer warps: chunk c is filled by producer c % np (one warp cannot keep up with the near field, whose
                                  // 9-row panels turn a chunk of 32 B rows into a handful of DMMAs per consumer warp)

__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }

struct BRing {
    unsigned char *base;
    uint64_t *full, *empty;
    uint32_t chunk_bytes, mask, log2n; // ring of 2^log2n chunks
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }

// The B producer warp: see above. Mirrors the consumers' walk (stages, runs, the applied-twice filter, the alignment of
// run / stage starts) so that both sides agree on the position of every column in the stream. One LANE per column works
// out where the column's row comes from; consecutive columns whose rows are consecutive in memory too — the columns of one
// low-rank piece (its T vectors are adjacent in the scratch, whose vector stride IS the row stride of the chunk), the
// columns of one dense leaf when the rows of the input matrix are VS doubles apart — are then copied by ONE bulk copy
// (TMA, completion counted in bytes on the chunk's mbarrier): ~4 copies per chunk of 32 rows instead of 32, and the warp
// never waits for data.
template <bool CPLX>
__device__ __forceinline__ void produce_b(const MSide &ks, const BlockDesc &bd, const MArgs &a, unsigned char *ring, uint32_t slot_bytes, uint64_t *full, uint64_t *empty, const BRing &br, uint32_t n_my_stages, int lane, uint32_t pid, uint32_t np) {
    const bool in16   = (a.ld_in % 2 == 0) && (a.col0_in % 2 == 0) && (a.mc % 2 == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 15u) == 0); // rows of the input matrix are 16 B aligned
    const bool in_seq = in16 && a.ld_in == a.vsp && a.col0_in == 0;                                                                        // ... and consecutive rows are one row of the chunk apart (the staged group)
    const uint32_t row_bytes = static_cast<uint32_t>(a.vsp) * 8u;
    RingPos pos;
    uint32_t bpos    = 0;  // position in the B stream
    long long open   = -1; // chunk being filled
    bool slow_copies = false; // the open chunk holds cp.async copies: wait for them before publishing
    auto publish     = [&]() { // hand the open chunk to the consumers: its phase completes when the bulk copies have landed
        if (open < 0)
            return;
        if (slow_copies) {
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            slow_copies = false;
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(smem_u32(&br.full[open & br.mask]));
        open = -1;
    };
    // b_global: the column tables are read from GLOBAL memory (the side's aux array) instead of the stage ring, so that the B rows
    // are fetched as far ahead as the B ring allows (8 chunks) and not only as far as the stage ring does (its 3 - 4 slots)
    uint32_t st_next = 0; // b_global: next stage of the block to look at
    for (uint32_t q = 0; q < n_my_stages; q++, pos.advance(ks.ring_stages)) {
        const unsigned char *auxrec;
        if (a.b_global) {
            StageDesc sd = ks.stages[bd.first_stage + st_next++];
            while (a.twice_only && !(sd.flags & 1u))
                sd = ks.stages[bd.first_stage + st_next++];
            auxrec = ks.aux + static_cast<size_t>(sd.aux_off16) * 16u;
        } else {
            mbar_wait(smem_u32(&full[pos.slot]), pos.phase);
            auxrec = ring + static_cast<size_t>(pos.slot) * slot_bytes + ks.stage_bytes;
        }
        const AuxHeader ah   = *reinterpret_cast<const AuxHeader *>(auxrec);
        const RunDesc *runs  = reinterpret_cast<const RunDesc *>(auxrec + sizeof(AuxHeader));
        const uint32_t *cols = reinterpret_cast<const uint32_t *>(runs + ah.n_runs);
        bpos                 = (bpos + 31u) & ~31u;
        for (uint32_t r = 0; r < ah.n_runs; r++) {
            const RunDesc rd = runs[r];
            if (a.twice_only && !(rd.flags & 1u))
                continue;
            bpos             = (bpos + 3u) & ~3u;
            const uint32_t K = a.skip_dense ? rd.K_lr : rd.K; // (near field applied from its own panels: the dense tail of the run is skipped)
            for (uint32_t j = 0; j < K;) {
                const long long chunk = static_cast<long long>((bpos + j) >> 5);
                const uint32_t first  = (bpos + j) & 31u;
                const uint32_t n      = (K - j) < (32u - first) ? (K - j) : (32u - first);
                if (static_cast<uint32_t>(chunk % np) != pid) { // another B producer's chunk
                    j += n;
                    continue;
                }
                if (chunk != open) {
                    publish();
                    if (lane == 0)
                        mbar_wait(smem_u32(&br.empty[chunk & br.mask]), static_cast<uint32_t>((chunk >> br.log2n) & 1) ^ 1u);
                    __syncwarp();
                    open = chunk;
                }
                const uint32_t bar  = smem_u32(&br.full[chunk & br.mask]);
                const uint32_t dst  = smem_u32(br.base + static_cast<size_t>(chunk & br.mask) * br.chunk_bytes) + (first + lane) * row_bytes;
                // lane c: where the row of column j + c comes from. key: rows with consecutive keys are consecutive in memory
                const double *p = nullptr;
                uint32_t mode = 3, key = 0xffffffffu; // 0 bulk copy, 1 cp.async 8 B pieces, 2 zero row, 3 nothing
                if (static_cast<uint32_t>(lane) < n) {
                    const uint32_t src = cols[rd.col0 + j + lane];
                    if (src & 0x80000000u) { // dense column: a row of the input matrix, mc doubles
                        const long long row = static_cast<long long>(src & 0x7fffffffu) + a.in_shift;
                        if (row < 0 || row >= a.in_rows)
                            mode = 2;
                        else {
                            p    = a.in + row * a.ld_in + a.col0_in;
                            mode = in16 ? 0u : 1u;
                            key  = in_seq ? src : 0xffffffffu;
                        }
                    } else { // a scratch vector (VS doubles, 16 B aligned), the next column of the piece is the next vector
                        p    = a.mscratch + static_cast<size_t>(src) * a.vsp;
                        mode = 0;
                        key  = src;
                    }
                }
                // heads of the sequences of consecutive rows; a head copies its whole sequence
                const uint32_t prev_key = __shfl_up_sync(0xffffffffu, key, 1);
                const bool cont         = lane > 0 && mode == 0 && key != 0xffffffffu && prev_key != 0xffffffffu && key == prev_key + 1u;
                const unsigned heads    = __ballot_sync(0xffffffffu, mode == 0 && !cont);
                const unsigned bulk     = __ballot_sync(0xffffffffu, mode == 0);
                const uint32_t one_row  = (mode == 0 && key == 0xffffffffu) ? static_cast<uint32_t>(a.mc) * 8u : row_bytes; // (a lone row of the input matrix: mc doubles)
                uint32_t my_bytes = 0;
                if (mode == 0 && !cont) {
                    const unsigned above = (lane == 31) ? 0u : ((heads | ~bulk) >> (lane + 1)); // next head, or the first lane that does not bulk-copy
                    const uint32_t len   = above ? static_cast<uint32_t>(__ffs(static_cast<int>(above))) : (32u - static_cast<uint32_t>(lane));
                    my_bytes             = len == 1 ? one_row : len * row_bytes;
                }
                // bytes of the batch: every bulk-copied row brings row_bytes, except lone rows of an input matrix narrower than VS
                const unsigned narrow = __ballot_sync(0xffffffffu, my_bytes != 0 && my_bytes < row_bytes);
                const uint32_t total  = static_cast<uint32_t>(__popc(bulk)) * row_bytes - static_cast<uint32_t>(__popc(narrow)) * (row_bytes - static_cast<uint32_t>(a.mc) * 8u);
                if (lane == 0 && total)
                    mbar_expect_tx(bar, total);
                __syncwarp();
                if (my_bytes)
                    bulk_g2s_plain(dst, p, my_bytes, bar);
                else if (mode == 1) {
                    for (int e = 0; e < a.mc; e++)
                        cp_async8(dst + 8u * e, p + e);
                } else if (mode == 2) {
                    for (int e = 0; e < a.vs; e++)
                        asm volatile("st.shared.f64 [%0], %1;" ::"r"(dst + 8u * e), "d"(0.) : "memory");
                }
                if (__any_sync(0xffffffffu, mode == 1))
                    slow_copies = true;
                j += n;
            }
            bpos += K;
        }
        publish(); // end of the stage: the consumers must not wait for the next stage to see its last chunk
        if (lane == 0 && !a.b_global)
            mbar_arrive(smem_u32(&empty[pos.slot])); // the column tables of this stage are no longer needed
    }
}

// A consumer warp enters chunk `chunk` of the B ring (releasing the one it was reading).
__device__ __forceinline__ void enter_chunk(const BRing &br, long long &cur, long long chunk, int lane) {
    if (chunk != cur) {
        if (cur >= 0) {
            __syncwarp();
            if (lane == 0)
                mbar_arrive(smem_u32(&br.empty[cur & br.mask]));
        }
        mbar_wait(smem_u32(&br.full[chunk & br.mask]), static_cast<uint32_t>((chunk >> br.log2n) & 1));
        cur = chunk;
    }
}

template <int NT, int NJ, int J0>
__device__ __forceinline__ void fold_tiles(double (&acc)[NJ][2], const double (&t)[NT][2]) {
#pragma unroll
    for (int jj = 0; jj < NT; jj++)
        if (J0 + jj < NJ) {
            acc[J0 + jj][0] += t[jj][0];
            acc[J0 + jj][1] += t[jj][1];
        }
}

// A run that meets NT <= 5 consecutive row tiles jlo .. jlo + NT - 1 of the warp's class (every run but the tall panels: the
// near field and the low-rank panels of the small clusters). The run is contracted into NT private accumulators with NO
// row predicates: a DMMA output row depends on the same row of A only, so a lane whose row lies outside the run
// may read whatever follows / precedes its panel in the ring slot (tile 0: clamped to row 0) — it pollutes rows of t that
// the fold below drops. The fold into the block accumulators (compile-time indices) happens once per run instead of a
// predicated walk over all NJ tiles at every k-step (legacy path 3: 20 - 35 instructions per DMMA; here 3 - 6). Columns
// >= mc of the B rows are not masked either: they only reach columns of C that the epilogue never stores.
template <bool CPLX, int NT, int PAR, int NJ>
__device__ __forceinline__ void run_tiles(const BRing &br, long long &cur, int lane, int tig, uint32_t run_pos, uint32_t Kr, const double *Prun, size_t pstep, size_t bstep,
                                          int vsp, int cBl, double bs, int rr0, int h, int jlo, double (&acc)[NJ][2]) {
    constexpr int CS = CPLX ? 1 : 0;
    constexpr int TS = (8 * PAR) << CS; // doubles between my rows of consecutive tiles of my class
    double t[NT][2], u[2] = {0., 0.};
#pragma unroll
    for (int jj = 0; jj < NT; jj++)
        t[jj][0] = t[jj][1] = 0.;
    const double *P0 = Prun + (static_cast<size_t>(rr0 > 0 ? rr0 : 0) << CS); // tile 0 (clamped)
    const double *PN = Prun + (static_cast<ptrdiff_t>(rr0) * (1 << CS));       // tiles >= 1 at PN[TS jj]: rr0 + 8 PAR jj > 0
    for (uint32_t k0 = 0; k0 < Kr;) {
        const uint32_t pos0 = run_pos + (k0 >> CS);
        const long long chunk = static_cast<long long>(pos0 >> 5);
        enter_chunk(br, cur, chunk, lane);
        uint32_t kend = k0 + ((32u - (pos0 & 31u)) << CS); // first contraction index of the next chunk
        if (kend > Kr)
            kend = Kr;
        const uint32_t kfull = (kend == Kr) ? (Kr & ~3u) : kend; // k-steps starting below kfull have four valid contraction indices
        const double *Bl = reinterpret_cast<const double *>(br.base + static_cast<size_t>(chunk & br.mask) * br.chunk_bytes) + static_cast<size_t>((pos0 & 31u) + (static_cast<uint32_t>(tig) >> CS)) * vsp + cBl;
        const double *A0 = P0 + static_cast<size_t>(k0 >> 2) * pstep;
        const double *AN = PN + static_cast<ptrdiff_t>(k0 >> 2) * static_cast<ptrdiff_t>(pstep);
        uint32_t k = k0;
        if (NT == 1) { // one tile: two accumulation chains over alternating k-steps hide the DMMA latency
            for (; k + 8 <= kfull; k += 8) {
                dmma(t[0], A0[0], CPLX ? bs * Bl[0] : Bl[0]);
                dmma(u, A0[pstep], CPLX ? bs * Bl[bstep] : Bl[bstep]);
                Bl += 2 * bstep;
                A0 += 2 * pstep;
            }
            if (k < kfull) {
                dmma(t[0], A0[0], CPLX ? bs * Bl[0] : Bl[0]);
                Bl += bstep;
                A0 += pstep;
                k += 4;
            }
            if (k < kend) { // last, partial k-step of the run: contraction indices >= Kr meet zeros on both sides
                const bool kv = k + tig < Kr;
                dmma(u, kv ? A0[0] : 0., kv ? bs * Bl[0] : 0.);
            }
        } else {
#pragma unroll 2
            for (; k < kfull; k += 4) {
                const double b = CPLX ? bs * Bl[0] : Bl[0];
                double av[NT]; // all the fragments of the k-step are in flight before its first DMMA
                av[0] = A0[0];
#pragma unroll
                for (int jj = 1; jj < NT; jj++)
                    av[jj] = AN[TS * jj];
#pragma unroll
                for (int jj = 0; jj < NT; jj++)
                    dmma(t[jj], av[jj], b);
                Bl += bstep;
                A0 += pstep;
                AN += pstep;
            }
            if (k < kend) {
                const bool kv  = k + tig < Kr;
                const double b = kv ? bs * Bl[0] : 0.;
                dmma(t[0], kv ? A0[0] : 0., b);
#pragma unroll
                for (int jj = 1; jj < NT; jj++)
                    dmma(t[jj], kv ? AN[TS * jj] : 0., b);
            }
        }
        k0 = kend;
    }
    if (NT == 1) {
        t[0][0] += u[0];
        t[0][1] += u[1];
    }
#pragma unroll
    for (int jj = 0; jj < NT; jj++) {
        const int rr  = rr0 + 8 * PAR * jj;
        const bool rv = rr >= 0 && rr < h;
        t[jj][0] = rv ? t[jj][0] : 0.;
        t[jj][1] = rv ? t[jj][1] : 0.;
    }
    // acc[jlo + jj] += t[jj] with compile-time indices (one case per jlo; a data-dependent index would push acc to local memory)
    switch (jlo) {
#define HTB_FOLD(J0)                                                                                                                                                 \
    case J0:                                                                                                                                                         \
        if constexpr (J0 < NJ)                                                                                                                                       \
            fold_tiles<NT, NJ, J0>(acc, t);                                                                                                                         \
        break;
        HTB_FOLD(0) HTB_FOLD(1) HTB_FOLD(2) HTB_FOLD(3) HTB_FOLD(4) HTB_FOLD(5) HTB_FOLD(6) HTB_FOLD(7)
#undef HTB_FOLD
    default: break;
    }
}

template <bool CPLX, int CT>
__global__ void __launch_bounds__((kApplyWarps + 1 + kBProducersMax) * 32, 1) apply_m_kernel(MSide ks, MArgs a, int b_log2n) {
    constexpr int CS  = CPLX ? 1 : 0;
    constexpr int PAR = kApplyWarps / CT;                            // row-tile classes
    constexpr int NJ  = (CPLX ? 8 : 16) / PAR > 0 ? (CPLX ? 8 : 16) / PAR : 1; // row tiles of a warp (a complex block has 64 rows = 8 tiles)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const BlockDesc bd = ks.blocks[ks.order[blockIdx.x]];
    if (a.twice_only && !(bd.flags & 1u))
        return;
    const uint32_t n_my_stages = a.twice_only ? bd.n_twice_stages : bd.n_stages;
    const uint32_t slot_bytes  = ks.stage_bytes + ks.aux_bytes;
    unsigned char *ring = smem_raw;
    BRing br;
    br.log2n       = static_cast<uint32_t>(b_log2n);
    br.mask        = (1u << b_log2n) - 1u;
    br.chunk_bytes = static_cast<uint32_t>(kBChunk) * static_cast<uint32_t>(a.vsp) * 8u;
    br.base        = smem_raw + static_cast<size_t>(ks.ring_stages) * slot_bytes;
    uint64_t *full  = reinterpret_cast<uint64_t *>(br.base + (static_cast<size_t>(kBChunk) * 72u * 8u << b_log2n));
    uint64_t *empty = full + ks.ring_stages;
    br.full         = empty + ks.ring_stages;
    br.empty        = br.full + (1u << b_log2n);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    const uint32_t np = (blockDim.x >> 5) - kApplyWarps - 1; // B producer warps

    if (threadIdx.x == 0) {
        for (int s = 0; s < ks.ring_stages; s++) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), kApplyWarps + (a.b_global ? 0u : np)); // the consumers and (unless they read the tables from global memory) the B producers
        }
        for (uint32_t s = 0; s <= br.mask; s++) {
            mbar_init(smem_u32(&br.full[s]), 1); // the B producer's arrive; the rows are counted in bytes (expect_tx)
            mbar_init(smem_u32(&br.empty[s]), kApplyWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (warp == kApplyWarps) {
        if (lane == 0)
            produce(ks, bd, ring, slot_bytes, full, empty, a.twice_only);
        return;
    }
    if (warp > kApplyWarps) {
        produce_b<CPLX>(ks, bd, a, ring, slot_bytes, full, empty, br, n_my_stages, lane, static_cast<uint32_t>(warp - kApplyWarps - 1), np);
        return;
    }
    const int ct = warp % CT, par = warp / CT;
    double acc[NJ][2];
#pragma unroll
    for (int j = 0; j < NJ; j++)
        acc[j][0] = acc[j][1] = 0.;
    // column of the B row this lane reads: c = 8 ct + g; complex, odd contraction index: the swapped partner with a sign
    const int cB       = 8 * ct + g;
    const bool c_valid = cB < a.mc;
    const int cBl      = (CPLX && (tig & 1)) ? (cB ^ 1) : cB; // (i v)[c] = c even ? -im(v) : re(v); conjugated panel: the opposite sign
    const double bs    = (CPLX && (tig & 1)) ? ((((cB & 1) != 0) == (a.conj == 0)) ? 1. : -1.) : 1.;
    const size_t bstep = static_cast<size_t>(4 >> CS) * a.vsp; // B rows per k-step
    const int my_row0  = 8 * par + g;                          // my row in tile j is my_row0 + 8 PAR j

    // Block constants of the FULL-HEIGHT fast path (below): most of the coefficients sit in stages made of ONE run over all the
    // rows of the block with <= 32 columns (a 24 KiB stage of a 122-row panel holds 25) — one chunk of the B ring, no row predicate.
    const uint32_t ld_blk    = CPLX ? 2u * static_cast<uint32_t>(bd.nrows) : unit_ld(static_cast<uint32_t>(bd.nrows), sizeof(double));
    const size_t pstep_blk   = static_cast<size_t>(4 >> CS) * ld_blk;
    const size_t poff_blk    = static_cast<size_t>(tig >> CS) * ld_blk + (CPLX ? (tig & 1) : 0) + (static_cast<size_t>(my_row0) << CS);
    const bool full_height   = a.fast_tall && ((bd.nrows - 1) >> 3) >= par + PAR * (NJ - 1); // my last tile meets rows of the block

    RingPos pos;
    uint32_t cpos = 0;  // position in the B stream (same walk as produce_b)
    long long cur = -1; // chunk the warp is reading
    for (uint32_t q = 0; q < n_my_stages; q++, pos.advance(ks.ring_stages)) {
        mbar_wait(smem_u32(&full[pos.slot]), pos.phase);
        const unsigned char *stage = ring + static_cast<size_t>(pos.slot) * slot_bytes;
        const StageHeader hdr      = *reinterpret_cast<const StageHeader *>(stage);
        const Unit *units          = reinterpret_cast<const Unit *>(stage + sizeof(StageHeader));
        const double *data         = reinterpret_cast<const double *>(stage + hdr.data_byte_off);
        const AuxHeader ah         = *reinterpret_cast<const AuxHeader *>(stage + ks.stage_bytes);
        const RunDesc *runs        = reinterpret_cast<const RunDesc *>(stage + ks.stage_bytes + sizeof(AuxHeader));
        cpos                       = (cpos + 31u) & ~31u;
        if (full_height && ah.n_runs == 1u && hdr.n_units == hdr.n_panel) {
            const RunDesc rd    = runs[0];
            const uint32_t Krun = a.skip_dense ? rd.K_lr : rd.K;
            if (rd.row0 == 0 && static_cast<int>(rd.h_minus_1) + 1 == bd.nrows && Krun != 0u && Krun <= 32u && !(a.twice_only && !(rd.flags & 1u))) {
                const long long chunk = static_cast<long long>(cpos >> 5);
                cpos += Krun;
                enter_chunk(br, cur, chunk, lane);
                const uint32_t Kr = Krun << CS, kfull = Kr & ~3u;
                const double *Bl  = reinterpret_cast<const double *>(br.base + static_cast<size_t>(chunk & br.mask) * br.chunk_bytes) + static_cast<size_t>(static_cast<uint32_t>(tig) >> CS) * a.vsp + cBl;
                const double *Pa  = data + (static_cast<size_t>(rd.data_off) << CS) + poff_blk;
                uint32_t k        = 0;
#pragma unroll 2
                for (; k < kfull; k += 4) {
                    const double b = CPLX ? bs * Bl[0] : Bl[0];
                    double av[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; j++)
                        av[j] = Pa[static_cast<size_t>(8 * PAR * j) << CS];
#pragma unroll
                    for (int j = 0; j < NJ; j++)
                        dmma(acc[j], av[j], b);
                    Bl += bstep;
                    Pa += pstep_blk;
                }
                if (k < Kr) { // last, partial k-step: contraction indices >= Kr meet zeros on both sides
                    const bool kv  = k + tig < Kr;
                    const double b = (c_valid && kv) ? bs * Bl[0] : 0.;
#pragma unroll
                    for (int j = 0; j < NJ; j++) {
                        const bool rv = 8 * PAR * j + my_row0 < bd.nrows;
                        dmma(acc[j], (kv && rv) ? Pa[static_cast<size_t>(8 * PAR * j) << CS] : 0., b);
                    }
                }
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(smem_u32(&empty[pos.slot]));
                continue;
            }
        }
        for (uint32_t r = 0; r < ah.n_runs; r++) {
            const RunDesc rd = runs[r];
            if (a.twice_only && !(rd.flags & 1u))
                continue;
            cpos = (cpos + 3u) & ~3u;
            const uint32_t run_pos = cpos;
            const uint32_t Krun    = a.skip_dense ? rd.K_lr : rd.K;
            if (Krun == 0)
                continue;
            cpos += Krun;
            const int row0 = rd.row0, h = static_cast<int>(rd.h_minus_1) + 1;
            // my row tiles j PAR + par that meet rows [row0, row0 + h): j in [jlo, jhi]
            const int tlo = row0 >> 3, thi = (row0 + h - 1) >> 3;
            const int jlo = tlo <= par ? 0 : (tlo - par + PAR - 1) / PAR, jhi = thi >= par ? (thi - par) / PAR : -1;
            const bool mine = jlo <= jhi && jlo < NJ; // (a warp without rows in the run still walks its chunks: the B ring is released by all)
            const uint32_t Kr = Krun << CS; // contraction length
            const uint32_t ld = CPLX ? 2u * static_cast<uint32_t>(h) : unit_ld(static_cast<uint32_t>(h), sizeof(double));
            const size_t pstep = static_cast<size_t>(4 >> CS) * ld;
            // this lane's panel column (contraction index tig) at my row of tile 0 — dereferenced only where the row exists
            const double *Prun = data + (static_cast<size_t>(rd.data_off) << CS) + static_cast<size_t>(tig >> CS) * ld + (CPLX ? (tig & 1) : 0);
            const bool tall    = row0 == 0 && h == bd.nrows && jlo == 0 && jhi == NJ - 1;
            if (mine && !tall && a.small_runs && jhi - jlo < 5) {
                const int rr0 = my_row0 + 8 * PAR * jlo - row0;
#define HTB_RUN_TILES(N)                                                                                                                                              \
    if constexpr (N <= NJ)                                                                                                                                           \
        run_tiles<CPLX, N, PAR, NJ>(br, cur, lane, tig, run_pos, Kr, Prun, pstep, bstep, a.vsp, cBl, bs, rr0, h, jlo, acc);
                switch (jhi - jlo) {
                case 0: HTB_RUN_TILES(1) break;
                case 1: HTB_RUN_TILES(2) break;
                case 2: HTB_RUN_TILES(3) break;
                case 3: HTB_RUN_TILES(4) break;
                default: HTB_RUN_TILES(5) break;
                }
#undef HTB_RUN_TILES
                continue;
            }
            const int path     = !mine ? 0 : (jlo == jhi ? 1 : (tall ? 2 : 3));
            // the k-steps of the run, chunk by chunk of the B ring
            for (uint32_t k0 = 0; k0 < Kr;) {
                const uint32_t pos0   = run_pos + (k0 >> CS);
                const long long chunk = static_cast<long long>(pos0 >> 5);
                if (chunk != cur) {
                    if (cur >= 0) {
                        __syncwarp();
                        if (lane == 0)
                            mbar_arrive(smem_u32(&br.empty[cur & br.mask]));
                    }
                    mbar_wait(smem_u32(&br.full[chunk & br.mask]), static_cast<uint32_t>((chunk >> br.log2n) & 1));
                    cur = chunk;
                }
                uint32_t kend = k0 + ((32u - (pos0 & 31u)) << CS); // first contraction index of the next chunk
                if (kend > Kr)
                    kend = Kr;
                const uint32_t kfull = (kend == Kr) ? (Kr & ~3u) : kend; // k-steps starting below kfull have four valid contraction indices
                if (path != 0) {
                    const double *Bl = reinterpret_cast<const double *>(br.base + static_cast<size_t>(chunk & br.mask) * br.chunk_bytes) + static_cast<size_t>((pos0 & 31u) + (static_cast<uint32_t>(tig) >> CS)) * a.vsp + cBl;
                    const double *Pl = Prun + static_cast<size_t>(k0 >> 2) * pstep;
                    uint32_t k       = k0;
                    if (path == 1) {
                        // ONE row tile (dense leaves, small clusters): two accumulation chains over alternating k-steps hide
                        // the DMMA latency; folded into the tile's accumulator at the end of the segment
                        const int rr   = my_row0 + 8 * PAR * jlo - row0;
                        const bool rv  = rr >= 0 && rr < h;
                        const double *Pa = Pl + (static_cast<size_t>(rv ? rr : 0) << CS);
                        double t0[2] = {0., 0.}, t1[2] = {0., 0.};
                        // (a lane supplies A[row g][k tig] and B[k tig][column g]: the two are valid independently)
                        for (; k + 8 <= kfull; k += 8) {
                            dmma(t0, rv ? Pa[0] : 0., c_valid ? bs * Bl[0] : 0.);
                            dmma(t1, rv ? Pa[pstep] : 0., c_valid ? bs * Bl[bstep] : 0.);
                            Bl += 2 * bstep;
                            Pa += 2 * pstep;
                        }
                        if (k < kfull) {
                            dmma(t0, rv ? Pa[0] : 0., c_valid ? bs * Bl[0] : 0.);
                            Bl += bstep;
                            Pa += pstep;
                            k += 4;
                        }
                        if (k < kend) { // last, partial k-step of the run
                            const bool kv = k + tig < Kr;
                            dmma(t1, (rv && kv) ? Pa[0] : 0., (c_valid && kv) ? bs * Bl[0] : 0.);
                        }
                        t0[0] += t1[0];
                        t0[1] += t1[1];
#pragma unroll
                        for (int j = 0; j < NJ; j++)
                            if (j == jlo) {
                                acc[j][0] += t0[0];
                                acc[j][1] += t0[1];
                            }
                    } else if (path == 2) {
                        // A run over ALL the rows of the block (the panels of the tall leaves: most of the coefficients): no
                        // row predicates. A lane whose row lies past the block reads the coefficient of a neighbouring
                        // column instead of a zero: it only pollutes accumulators of rows >= nrows, which are never stored.
                        const double *Pa = Pl + (static_cast<size_t>(my_row0) << CS);
#pragma unroll 2
                        for (; k < kfull; k += 4) {
                            const double b = CPLX ? bs * Bl[0] : Bl[0]; // (columns >= mc: garbage that only reaches columns of C never stored)
                            double av[NJ]; // all the fragments of the k-step are in flight before its first DMMA
#pragma unroll
                            for (int j = 0; j < NJ; j++)
                                av[j] = Pa[static_cast<size_t>(8 * PAR * j) << CS];
#pragma unroll
                            for (int j = 0; j < NJ; j++)
                                dmma(acc[j], av[j], b);
                            Bl += bstep;
                            Pa += pstep;
                        }
                        if (k < kend) { // last, partial k-step of the run: contraction indices >= Kr meet zeros on both sides
                            const bool kv  = k + tig < Kr;
                            const double b = (c_valid && kv) ? bs * Bl[0] : 0.;
#pragma unroll
                            for (int j = 0; j < NJ; j++) {
                                const bool rv = 8 * PAR * j + my_row0 < h;
                                dmma(acc[j], (kv && rv) ? Pa[static_cast<size_t>(8 * PAR * j) << CS] : 0., b);
                            }
                        }
                    } else {
                        for (; k < kend; k += 4) {
                            const bool kv  = k + tig < Kr;
                            const double b = (c_valid && kv) ? bs * Bl[0] : 0.;
#pragma unroll
                            for (int j = 0; j < NJ; j++) {
                                if (j >= jlo && j <= jhi) { // warp-uniform
                                    const int rr    = my_row0 + 8 * PAR * j - row0;
                                    const bool rv   = rr >= 0 && rr < h;
                                    const double av = (rv && kv) ? Pl[static_cast<size_t>(rv ? rr : 0) << CS] : 0.;
                                    dmma(acc[j], av, b);
                                }
                            }
                            Bl += bstep;
                            Pl += pstep;
                        }
                    }
                }
                k0 = kend;
            }
        }
        // ADDVEC units (side 1, transposed direction): C rows += the TF vectors REDUCE_M produced for dense leaves
        if (hdr.n_units > hdr.n_panel) {
            const MUnit *mun = ks.munits + hdr.first_unit;
            for (uint32_t u = hdr.n_panel; u < hdr.n_units; u++) {
                const Unit un = units[u];
                if (a.twice_only && !unit_twice(un.geom))
                    continue;
                const int row0 = static_cast<int>(unit_row0(un.geom)), h = static_cast<int>(unit_h(un.geom));
                const uint32_t src = mun[u].src;
                const int c        = 8 * ct + 2 * tig;
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const int rr = my_row0 + 8 * PAR * j - row0;
                    if (rr >= 0 && rr < h && c < a.vs) {
                        const double2 v = *reinterpret_cast<const double2 *>(a.mscratch + static_cast<size_t>(src + rr) * a.vsp + c);
                        acc[j][0] += v.x;
                        acc[j][1] += v.y;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(smem_u32(&empty[pos.slot]));
    }
    // epilogue: alpha / beta, one write per C entry (complex: a lane holds re and im of one entry)
    const int c = 8 * ct + 2 * tig;
#pragma unroll
    for (int j = 0; j < NJ; j++) {
        const int i = my_row0 + 8 * PAR * j;
        if (i < bd.nrows && c < a.mc) {
            const long long gr = static_cast<long long>(bd.row_start) + i + a.out_shift;
            if (gr >= 0 && gr < a.out_rows) {
                double *o = a.out + gr * a.ld_out + a.col0 + c;
                if (CPLX) {
                    double re = a.alpha * acc[j][0] - a.alpha_im * acc[j][1], im = a.alpha * acc[j][1] + a.alpha_im * acc[j][0];
                    if (!a.beta_is_zero) {
                        const double ore = o[0], oim = o[1];
                        re += a.beta * ore - a.beta_im * oim;
                        im += a.beta * oim + a.beta_im * ore;
                    }
                    o[0] = re;
                    o[1] = im;
                } else {
#pragma unroll
                    for (int jj = 0; jj < 2; jj++)
                        if (c + jj < a.mc) {
                            double r = a.alpha * acc[j][jj];
                            if (!a.beta_is_zero)
                                r = fma(a.beta, o[jj], r);
                            o[jj] = r;
                        }
                }
            }
        }
    }
}

// One warp per (piece, vector of the piece): sums the per-chunk partial vectors in chunk order (fixed summation order).
// A lane owns two adjacent columns (one 16 B load per partial); the loads of 8 partials are in flight together, the adds
// keep their order.
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
    const size_t step = static_cast<size_t>(len) * vsp;
    for (uint32_t c = 2 * lane; c < static_cast<uint32_t>(vs); c += 64) {
        const double *p = mscratch + (static_cast<size_t>(ce.src) + k) * vsp + c;
        double2 v       = make_double2(0., 0.);
        uint32_t j      = 0;
        for (; j + 8 <= n_sum; j += 8) {
            double2 t[8];
#pragma unroll
            for (int q = 0; q < 8; q++)
                t[q] = *reinterpret_cast<const double2 *>(p + static_cast<size_t>(j + q) * step);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                v.x += t[q].x;
                v.y += t[q].y;
            }
        }
        for (; j < n_sum; j++) {
            const double2 t = *reinterpret_cast<const double2 *>(p + static_cast<size_t>(j) * step);
            v.x += t.x;
            v.y += t.y;
        }
        *reinterpret_cast<double2 *>(mscratch + (static_cast<size_t>(ce.dst_first) + k) * vsp + c) = v;
    }
}

__global__ void stage_group_kernel(const double *__restrict__ in, long long rows, int ld_in, int col0, int mc, double *__restrict__ dst, int vsp) {
    const long long total = rows * vsp;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = idx / vsp;
        const int c       = static_cast<int>(idx - r * vsp);
        dst[idx]          = c < mc ? in[r * ld_in + col0 + c] : 0.;
    }
}

// One warp per task: an h x w panel of doubles (complex: 2 h real rows) from the main stream into a near-field panel.
__global__ void nf_copy_kernel(const NfTask *tasks, long long n_tasks, const unsigned char *src, unsigned char *dst, int real_per_elem) {
    const long long t = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tasks)
        return;
    const NfTask task = tasks[t];
    const int lane    = threadIdx.x & 31;
    const double *s   = reinterpret_cast<const double *>(src + task.src_off);
    double *d         = reinterpret_cast<double *>(dst + task.dst_off);
    const int h = task.h * real_per_elem, sld = task.src_ld * real_per_elem, dld = task.dst_ld * real_per_elem;
    const int total = h * task.w;
    for (int e = lane; e < total; e += 32) {
        const int k = e / h, i = e - k * h;
        atomicAdd(&d[static_cast<size_t>(k) * dld + i], s[static_cast<size_t>(k) * sld + i]); // (leaves of a block tree never overlap: 0 + x, exact; overlapping leaf lists add up)
    }
}

inline size_t aux_part(const LaunchConfig &cfg) { return cfg.m_aux_bytes > 0 ? static_cast<size_t>(cfg.m_aux_bytes) : aux_slot_bytes(static_cast<uint32_t>(cfg.cseg_bytes)); }

inline MSide make_mside(const SideDevice &s, const LaunchConfig &cfg, int ring, bool apply_role) {
    return MSide{s.blocks, s.stages, s.order, s.stream, apply_role ? s.aux_apply : s.aux_reduce, s.munits, cfg.m_x_rows > 0 ? cfg.m_x_rows : cfg.block_rows, cfg.stage_bytes, static_cast<int>(aux_part(cfg)), ring};
}

} // namespace

size_t reduce_m_smem_bytes(const LaunchConfig &cfg, int vs, size_t esize) {
    const size_t slot = static_cast<size_t>(cfg.stage_bytes) + aux_part(cfg);
    const size_t rb   = static_cast<size_t>(cfg.m_x_rows > 0 ? cfg.m_x_rows : cfg.block_rows) * (esize / 8);
    return static_cast<size_t>(cfg.m_reduce_ring_stages) * slot + sizeof(double) * (rb + 4) * (vs + cfg.m_pad) + 16 * static_cast<size_t>(cfg.m_reduce_ring_stages) + 16;
}
size_t apply_m_smem_bytes(const LaunchConfig &cfg) {
    const size_t slot = static_cast<size_t>(cfg.stage_bytes) + aux_part(cfg);
    const size_t nb   = size_t(1) << cfg.m_b_ring_log2;
    return static_cast<size_t>(cfg.m_ring_stages) * slot + nb * kBChunk * 72 * 8 + 16 * static_cast<size_t>(cfg.m_ring_stages) + 16 * nb;
}

cudaError_t configure_mkernels(const LaunchConfig &cfg, size_t esize) {
    // The attribute belongs to the kernel, not to the handle, and ring slots are sized per store (largest aux record): two
    // operators of one process need different amounts. Allow whatever the device offers; each launch passes its own size.
    int dev = 0, optin = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 == cudaSuccess)
        e0 = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e0 != cudaSuccess)
        return e0;
    auto set = [optin](const void *f, size_t smem) -> cudaError_t {
        if (smem > static_cast<size_t>(optin))
            return cudaErrorInvalidValue;
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
        // several CTAs per SM: ask for the largest shared-memory carve-out, the default only guarantees one block
        return e != cudaSuccess ? e : cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    };
    cudaError_t e;
    const void *red = esize == 16 ? reinterpret_cast<const void *>(reduce_m_kernel<true>) : reinterpret_cast<const void *>(reduce_m_kernel<false>);
    if ((e = set(red, reduce_m_smem_bytes(cfg, 64, esize))) != cudaSuccess)
        return e;
    const void *app[4];
    if (esize == 16) {
        app[0] = reinterpret_cast<const void *>(apply_m_kernel<true, 1>), app[1] = reinterpret_cast<const void *>(apply_m_kernel<true, 2>);
        app[2] = reinterpret_cast<const void *>(apply_m_kernel<true, 4>), app[3] = reinterpret_cast<const void *>(apply_m_kernel<true, 8>);
    } else {
        app[0] = reinterpret_cast<const void *>(apply_m_kernel<false, 1>), app[1] = reinterpret_cast<const void *>(apply_m_kernel<false, 2>);
        app[2] = reinterpret_cast<const void *>(apply_m_kernel<false, 4>), app[3] = reinterpret_cast<const void *>(apply_m_kernel<false, 8>);
    }
    for (const void *f : app)
        if ((e = set(f, apply_m_smem_bytes(cfg))) != cudaSuccess)
            return e;
    return cudaSuccess;
}

cudaError_t launch_reduce_m(const SideDevice &side, const LaunchConfig &cfg, const MArgs &args, cudaStream_t stream) {
    if (side.n_blocks == 0)
        return cudaSuccess;
    const MSide ms    = make_mside(side, cfg, cfg.m_reduce_ring_stages, false);
    const int threads = (std::min(std::max(cfg.m_reduce_warps, 4), kReduceWarpsMax) + 1) * 32;
    if (args.cplx)
        reduce_m_kernel<true><<<side.n_blocks, threads, reduce_m_smem_bytes(cfg, args.vs, 16), stream>>>(ms, args);
    else
        reduce_m_kernel<false><<<side.n_blocks, threads, reduce_m_smem_bytes(cfg, args.vs, 8), stream>>>(ms, args);
    return cudaGetLastError();
}

cudaError_t launch_apply_m(const SideDevice &side, const LaunchConfig &cfg, const MArgs &args, cudaStream_t stream) {
    if (side.n_blocks == 0)
        return cudaSuccess;
    MArgs a            = args;
    a.beta_is_zero     = (args.beta == 0. && args.beta_im == 0.) ? 1 : 0;
    const MSide ms     = make_mside(side, cfg, cfg.m_ring_stages, true);
    const size_t smem  = apply_m_smem_bytes(cfg);
    const int threads  = (kApplyWarps + 1 + std::min(std::max(cfg.m_b_producers, 1), kBProducersMax)) * 32;
    const int bl       = cfg.m_b_ring_log2;
    // column tiles of the group: no warp is given zero columns to multiply
    const int ct = args.vs > 32 ? 8 : (args.vs > 16 ? 4 : (args.vs > 8 ? 2 : 1));
#define HTB_APPLY_M(C, T) apply_m_kernel<C, T><<<side.n_blocks, threads, smem, stream>>>(ms, a, bl)
    if (args.cplx) {
        if (ct == 8)
            HTB_APPLY_M(true, 8);
        else if (ct == 4)
            HTB_APPLY_M(true, 4);
        else if (ct == 2)
            HTB_APPLY_M(true, 2);
        else
            HTB_APPLY_M(true, 1);
    } else {
        if (ct == 8)
            HTB_APPLY_M(false, 8);
        else if (ct == 4)
            HTB_APPLY_M(false, 4);
        else if (ct == 2)
            HTB_APPLY_M(false, 2);
        else
            HTB_APPLY_M(false, 1);
    }
#undef HTB_APPLY_M
    return cudaGetLastError();
}

cudaError_t launch_nf_copy(const NfTask *tasks, long long n_tasks, const unsigned char *src, unsigned char *dst, int esize, cudaStream_t stream) {
    if (n_tasks <= 0)
        return cudaSuccess;
    const int threads   = 256;
    const unsigned grid = static_cast<unsigned>((n_tasks * 32 + threads - 1) / threads);
    nf_copy_kernel<<<grid, threads, 0, stream>>>(tasks, n_tasks, src, dst, esize / 8);
    return cudaGetLastError();
}

cudaError_t launch_stage_group(const double *in, long long rows, int ld_in, int col0, int mc, double *dst, int vsp, cudaStream_t stream) {
    if (rows <= 0)
        return cudaSuccess;
    const long long total = rows * vsp;
    const unsigned grid   = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148LL * 32));
    stage_group_kernel<<<grid, 256, 0, stream>>>(in, rows, ld_in, col0, mc, dst, vsp);
    return cudaGetLastError();
}

cudaError_t launch_combine_m(const SideDevice &side, double *mscratch, int vs, int vsp, int twice_only, cudaStream_t stream) {
    if (side.n_combine_m == 0)
        return cudaSuccess;
    const int warps = 8; // 32 warps (one per vector of the piece, pieces hold <= 32 vectors) per entry
    combine_m_kernel<<<static_cast<unsigned>((static_cast<long long>(side.n_combine_m) * 32 + warps - 1) / warps), warps * 32, 0, stream>>>(side.combine_m, side.n_combine_m, mscratch, vs, vsp, twice_only);
    return cudaGetLastError();
}

} // namespace htb
