// htool_b200/csrc/kernels.cu — sm_100a kernels of the H-matrix product.
//
// What they replace in the reference (CPU, OpenMP + BLAS):
//   * the leaf loop of openmp_internal_add_hmatrix_vector_product
//     (include/htool/hmatrix/linalg/add_hmatrix_vector_product.hpp:139-169), i.e. one Blas::gemv per dense
//     leaf (matrix/linalg/add_matrix_vector_product.hpp:10-18) and two per low-rank leaf
//     (hmatrix/lrmat/linalg/add_lrmat_vector_product.hpp:9-24), a heap temp per leaf, per-thread output
//     copies and a serialised axpy reduction.
// How: every CTA owns one BLOCK of a side (store.hpp) and consumes that block's stream, stage by stage,
// through a shared-memory ring filled by 1-D bulk-async copies (cp.async.bulk, the TMA engine; SASS
// UBLKCP) that a single producer lane issues and mbarriers track. Eight consumer warps take the units of
// the stream round-robin (one unit = one h x w column-major panel, h <= block_rows, w <= 32):
//   REDUCE  t[k] = sum_i op(P[i,k]) x[i]: the block's x sub-vector is staged in shared memory once. Lanes run
//           along the rows (several columns side by side when h <= 16); the per-lane products of up to 8
//           columns are folded with a TRANSPOSING butterfly (7 + log2(seg) - 3 shuffles for 8 columns
//           instead of 8 * log2(seg)), in a fixed order;
//   APPLY   y[i] += sum_k op(P[i,k]) c[k]: the c vectors of a stage (t pieces, x slices) arrive in shared
//           memory with the stage itself (the c-stream segment, a second bulk copy on the same mbarrier).
//           Lanes run along the rows, per-warp private y accumulators live in shared memory, are summed
//           over the warps in warp order at the end and written once (y rows are owned by exactly one
//           CTA): deterministic, no atomics, beta/alpha fused in the same epilogue.
// Bandwidth-bound: every coefficient crosses HBM->SMEM once per pass and is used for one FMA.
#include "kernels.cuh"

#include <cstdint>

namespace htb {

bool g_pdl = true; // programmatic dependent launch of the pass kernels (option "pdl", read at htb_create)

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kThreads       = (kConsumerWarps + 1) * 32; // + 1 producer warp
constexpr int kMaxQ          = 2;                         // a lane owns at most 2 row slabs (block_rows * sizeof(T) <= 1024)

// ---- scalar helpers -------------------------------------------------------------------------------
__device__ __forceinline__ double zero_of(double) { return 0.; }
__device__ __forceinline__ cplx zero_of(cplx) { return cplx{0., 0.}; }
__device__ __forceinline__ double add(double a, double b) { return a + b; }
__device__ __forceinline__ cplx add(cplx a, cplx b) { return cplx{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ double mul(double a, double b) { return a * b; }
__device__ __forceinline__ cplx mul(cplx a, cplx b) { return cplx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
// c + a*b
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ cplx fma_(cplx a, cplx b, cplx c) {
    return cplx{fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y))};
}
template <bool CONJ>
__device__ __forceinline__ double cj(double a) { return a; }
template <bool CONJ>
__device__ __forceinline__ cplx cj(cplx a) { return CONJ ? cplx{a.x, -a.y} : a; }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ cplx shfl_xor(cplx v, int m) { return cplx{__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m)}; }
__device__ __forceinline__ double select(bool p, double a, double b) { return p ? a : b; }
__device__ __forceinline__ cplx select(bool p, cplx a, cplx b) { return cplx{p ? a.x : b.x, p ? a.y : b.y}; }

// R consecutive rows of one column are fetched with ONE 128-bit shared-memory load: 2 doubles or 1 complex
template <typename T>
struct Rows;
template <>
struct Rows<double> {
    static constexpr int R = 2;
};
template <>
struct Rows<cplx> {
    static constexpr int R = 1;
};
__device__ __forceinline__ void load_rows(const double *p, double (&f)[2]) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    f[0]            = v.x;
    f[1]            = v.y;
}
__device__ __forceinline__ void load_rows(const cplx *p, cplx (&f)[1]) { f[0] = *p; }

// ---- mbarrier / bulk copy (inline PTX) -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait suspends the thread in hardware up to the time hint, so a waiting warp costs few issue slots
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
// global -> shared bulk async copy, completion counted in bytes on an mbarrier (TMA engine, 1-D)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy, bool hint) {
    if (hint)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                     : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// barrier among the consumer warps only (the producer warp is busy streaming)
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory"); }

// programmatic dependent launch (see launch_pdl): let the successor be scheduled / wait for the predecessor's results
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct KernelSide {
    const BlockDesc *blocks;
    const StageDesc *stages;
    const uint32_t *order;
    const unsigned char *stream;
    unsigned long long cs_base;
    int block_rows, stage_bytes, cseg_bytes, ring_stages, evict_first;
};

// Shared-memory carve-up common to both kernels: [ring (slot = stage [+ c segment]) | vec | barriers]
struct SmemLayout {
    unsigned char *ring;
    unsigned char *vec; // REDUCE: x sub-vector (block_rows). APPLY: per-warp y accumulators (warps x block_rows)
    uint64_t *full, *empty;
    uint32_t slot_bytes;
};
__device__ __forceinline__ SmemLayout carve(unsigned char *base, const KernelSide &ks, uint32_t slot_bytes, size_t vec_bytes) {
    SmemLayout s;
    s.ring       = base;
    s.slot_bytes = slot_bytes;
    s.vec        = base + static_cast<size_t>(ks.ring_stages) * slot_bytes;
    s.full       = reinterpret_cast<uint64_t *>(s.vec + vec_bytes);
    s.empty      = s.full + ks.ring_stages;
    return s;
}

// Position of consecutive stages in the shared-memory ring: slot and phase parity, advanced without divisions.
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

// Producer: one lane streams the block's stages (and, for APPLY, their c segments) into the ring.
template <typename T, bool WITH_C>
__device__ __forceinline__ void produce(const KernelSide &ks, const BlockDesc &bd, const SmemLayout &sm, int twice_only, const T *cs) {
    const uint64_t policy = ks.evict_first ? l2_evict_first_policy() : 0;
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
        const uint32_t cbytes = WITH_C ? static_cast<uint32_t>(sd.c_len * sizeof(T)) : 0u;
        const uint32_t slot   = pos.slot;
        const uint32_t full   = smem_u32(&sm.full[slot]);
        mbar_wait(smem_u32(&sm.empty[slot]), pos.phase ^ 1u);
        mbar_arrive_expect_tx(full, sd.nbytes + cbytes);
        const uint32_t dst = smem_u32(sm.ring + static_cast<size_t>(slot) * sm.slot_bytes);
        bulk_g2s(dst, ks.stream + sd.byte_off, sd.nbytes, full, policy, ks.evict_first != 0);
        if (WITH_C && cbytes)
            bulk_g2s(dst + ks.stage_bytes, cs + sd.c_off, cbytes, full, 0, false);
        pos.advance(ks.ring_stages);
    }
}

__device__ __forceinline__ void init_barriers(const KernelSide &ks, const SmemLayout &sm) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < ks.ring_stages; s++) {
            mbar_init(smem_u32(&sm.full[s]), 1);               // the producer's arrive.expect_tx
            mbar_init(smem_u32(&sm.empty[s]), kConsumerWarps); // every consumer warp walks every stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

// Lane geometry of a unit of height h. A lane owns R consecutive rows (Rows<T>::R); hp = ceil(h / R) lane-rows are
// needed. Lanes are grouped in segments of seg = 2^seglog >= min(hp, 32) lanes: lane li of a segment owns rows
// R*li .. R*li + R - 1 (and R*(li + 32) .. when hp > 32) and the G = 32 / seg segments work on different columns.
struct LaneMap {
    int seglog, li, g, logG, Q;
};
template <int R>
__device__ __forceinline__ LaneMap lane_map(uint32_t h, int lane) {
    const uint32_t hp = (h + R - 1) / R;
    LaneMap m;
    m.seglog = hp > 16 ? 5 : (hp > 8 ? 4 : (hp > 4 ? 3 : (hp > 2 ? 2 : (hp > 1 ? 1 : 0))));
    m.li     = lane & ((1 << m.seglog) - 1);
    m.g      = lane >> m.seglog;
    m.logG   = 5 - m.seglog;
    m.Q      = hp > 32 ? 2 : 1;
    return m;
}

// v[c], c < J, per lane. For every c, sums v[c] over the lanes of each segment in a fixed order. On return the
// first nv entries of v hold the totals of indices cbase .. cbase + nv - 1, and all lanes that agree on the bits
// above lowmask hold the same totals. Halving steps exchange half of the values, so J values cost J - 1 (+ the
// remaining butterfly steps) shuffles instead of J * log2(seg).
template <typename T, int J>
__device__ __forceinline__ void seg_reduce(T (&v)[J], int seglog, int li, int &cbase, int &nv, int &lowmask) {
    int d     = (1 << seglog) >> 1;
    int lastd = 1 << seglog;
    cbase     = 0;
    nv        = J;
#pragma unroll
    for (int n = J / 2; n >= 1; n /= 2) {
        if (d >= 1) {
            const bool up = (li & d) != 0;
#pragma unroll
            for (int c = 0; c < n; c++) {
                const T keep = select(up, v[c + n], v[c]);
                const T send = select(up, v[c], v[c + n]);
                v[c]         = add(keep, shfl_xor(send, d));
            }
            cbase += up ? n : 0;
            nv    = n;
            lastd = d;
            d >>= 1;
        }
    }
    while (d >= 1) {
        v[0] = add(v[0], shfl_xor(v[0], d));
        d >>= 1;
    }
    lowmask = lastd - 1;
}

// Columns kb + (c << logG) + g, c < J, of the panel: per-lane products with x, segment reduction, store of the totals.
template <typename T, bool CONJ, int J>
__device__ __forceinline__ void reduce_batch(const T *P, uint32_t ld, uint32_t w, uint32_t kb, const LaneMap &m, const uint32_t (&off)[kMaxQ], const T (&xv)[kMaxQ][Rows<T>::R], T *out) {
    constexpr int R = Rows<T>::R;
    T v[J];
#pragma unroll
    for (int c = 0; c < J; c++) {
        const uint32_t k = kb + (c << m.logG) + m.g;
        const T *col     = P + (k < w ? k : w - 1) * ld; // clamped: the total of a column >= w is never stored
        T f[R];
        load_rows(col + off[0], f);
        T s = mul(cj<CONJ>(f[0]), xv[0][0]);
        if (R > 1)
            s = fma_(cj<CONJ>(f[R - 1]), xv[0][R - 1], s);
        if (m.Q > 1) {
            load_rows(col + off[1], f);
            s = fma_(cj<CONJ>(f[0]), xv[1][0], s);
            if (R > 1)
                s = fma_(cj<CONJ>(f[R - 1]), xv[1][R - 1], s);
        }
        v[c] = s;
    }
    int cbase, nv, lowmask;
    seg_reduce<T, J>(v, m.seglog, m.li, cbase, nv, lowmask);
    if ((m.li & lowmask) == 0) {
        if (nv == 1) { // the usual case: one total per writer lane
            const uint32_t k = kb + (cbase << m.logG) + m.g;
            if (k < w)
                out[k] = v[0];
        } else {
#pragma unroll
            for (int c = 0; c < J; c++) {
                const uint32_t k = kb + ((cbase + c) << m.logG) + m.g;
                if (c < nv && k < w)
                    out[k] = v[c];
            }
        }
    }
}

// REDUCE of one coefficient-carrying unit: out[k] = sum_i op(P[i,k]) xin[row0 + i], k < w
template <typename T, bool CONJ>
__device__ __forceinline__ void reduce_unit(const Unit &un, const T *data, const T *xin, T *scratch, int lane) {
    constexpr int R = Rows<T>::R;
    const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom), w = unit_w(un.geom);
    T *out            = scratch + un.out;
    const T *P        = data + un.data_off;
    const uint32_t ld = unit_ld(h, sizeof(T));
    const LaneMap m   = lane_map<R>(h, lane);
    uint32_t off[kMaxQ];
    T xv[kMaxQ][R];
#pragma unroll
    for (int qq = 0; qq < kMaxQ; qq++) {
        const uint32_t i0 = R * (m.li + 32u * qq);
        off[qq]           = i0 < ld ? i0 : 0u;
#pragma unroll
        for (int r = 0; r < R; r++)
            xv[qq][r] = i0 + r < h ? xin[row0 + i0 + r] : zero_of(T{});
    }
    for (uint32_t kb = 0; kb < w;) {
        const uint32_t per_seg = (w - kb + (1u << m.logG) - 1u) >> m.logG; // columns left for each segment
        if (sizeof(T) == 8 && per_seg > 4) { // (complex: batches of 4 at most, 8 complex partial sums cost an occupancy step)
            reduce_batch<T, CONJ, 8>(P, ld, w, kb, m, off, xv, out);
            kb += 8u << m.logG;
        } else if (per_seg > 2) {
            reduce_batch<T, CONJ, 4>(P, ld, w, kb, m, off, xv, out);
            kb += 4u << m.logG;
        } else {
            reduce_batch<T, CONJ, 2>(P, ld, w, kb, m, off, xv, out);
            kb += 2u << m.logG;
        }
    }
}

// ---- fused APPLY + REDUCE of one unit (symmetric storage, second application) ---------------------------------------
// One walk over the panel serves both products: every 128-bit shared-memory load of P feeds y[i] += op(P[i,k]) c[k]
// (accumulated per lane in acc) AND t'[k] = sum_i op2(P[i,k]) x2[i] (per-lane products folded by the transposing
// butterfly). Same lane / segment geometry as reduce_batch; column k belongs to segment k mod G in both.
template <typename T, bool CONJ, bool CONJ2, int J>
__device__ __forceinline__ void fused_batch(const T *P, uint32_t ld, uint32_t w, uint32_t kb, const LaneMap &m, const uint32_t (&off)[kMaxQ], const T (&xv)[kMaxQ][Rows<T>::R],
                                            const T *c, T (&acc)[kMaxQ][Rows<T>::R], T *out) {
    constexpr int R = Rows<T>::R;
    T v[J];
#pragma unroll
    for (int cc = 0; cc < J; cc++) {
        const uint32_t k  = kb + (cc << m.logG) + m.g;
        const bool valid  = k < w;
        const uint32_t kc = valid ? k : w - 1;
        const T *col      = P + kc * ld;
        const T ck        = valid ? c[kc] : zero_of(T{});
        T f[R];
        load_rows(col + off[0], f);
        T s = mul(cj<CONJ2>(f[0]), xv[0][0]);
        if (R > 1)
            s = fma_(cj<CONJ2>(f[R - 1]), xv[0][R - 1], s);
#pragma unroll
        for (int r = 0; r < R; r++)
            acc[0][r] = fma_(cj<CONJ>(f[r]), ck, acc[0][r]);
        if (m.Q > 1) {
            load_rows(col + off[1], f);
            s = fma_(cj<CONJ2>(f[0]), xv[1][0], s);
            if (R > 1)
                s = fma_(cj<CONJ2>(f[R - 1]), xv[1][R - 1], s);
#pragma unroll
            for (int r = 0; r < R; r++)
                acc[1][r] = fma_(cj<CONJ>(f[r]), ck, acc[1][r]);
        }
        v[cc] = s;
    }
    int cbase, nv, lowmask;
    seg_reduce<T, J>(v, m.seglog, m.li, cbase, nv, lowmask);
    if ((m.li & lowmask) == 0) {
        if (nv == 1) {
            const uint32_t k = kb + (cbase << m.logG) + m.g;
            if (k < w)
                out[k] = v[0];
        } else {
#pragma unroll
            for (int cc = 0; cc < J; cc++) {
                const uint32_t k = kb + ((cbase + cc) << m.logG) + m.g;
                if (cc < nv && k < w)
                    out[k] = v[cc];
            }
        }
    }
}

template <typename T, bool CONJ, bool CONJ2>
__device__ __forceinline__ void fused_unit(const Unit &un, const T *data, const T *c, const T *xin, T *yacc, T *scratch2, int lane) {
    constexpr int R = Rows<T>::R;
    const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom), w = unit_w(un.geom);
    T *out            = scratch2 + un.out;
    const T *P        = data + un.data_off;
    const uint32_t ld = unit_ld(h, sizeof(T));
    const LaneMap m   = lane_map<R>(h, lane);
    uint32_t off[kMaxQ];
    T xv[kMaxQ][R], acc[kMaxQ][R];
#pragma unroll
    for (int qq = 0; qq < kMaxQ; qq++) {
        const uint32_t i0 = R * (m.li + 32u * qq);
        off[qq]           = i0 < ld ? i0 : 0u;
#pragma unroll
        for (int r = 0; r < R; r++) {
            xv[qq][r]  = i0 + r < h ? xin[row0 + i0 + r] : zero_of(T{});
            acc[qq][r] = zero_of(T{});
        }
    }
    for (uint32_t kb = 0; kb < w;) {
        const uint32_t per_seg = (w - kb + (1u << m.logG) - 1u) >> m.logG;
        if (sizeof(T) == 8 && per_seg > 4) {
            fused_batch<T, CONJ, CONJ2, 8>(P, ld, w, kb, m, off, xv, c, acc, out);
            kb += 8u << m.logG;
        } else if (per_seg > 2) {
            fused_batch<T, CONJ, CONJ2, 4>(P, ld, w, kb, m, off, xv, c, acc, out);
            kb += 4u << m.logG;
        } else {
            fused_batch<T, CONJ, CONJ2, 2>(P, ld, w, kb, m, off, xv, c, acc, out);
            kb += 2u << m.logG;
        }
    }
    // y: the G segments worked on interleaved columns, fold them, then one lane per row adds into the warp's accumulator
    for (int d = 1 << m.seglog; d < 32; d <<= 1)
#pragma unroll
        for (int r = 0; r < R; r++)
            acc[0][r] = add(acc[0][r], shfl_xor(acc[0][r], d));
    if (m.g == 0) {
#pragma unroll
        for (int qq = 0; qq < kMaxQ; qq++)
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint32_t i = R * (m.li + 32u * qq) + r;
                if (i < h && (qq == 0 || m.Q > 1))
                    yacc[row0 + i] = add(yacc[row0 + i], acc[qq][r]);
            }
    }
}

// ---- REDUCE -------------------------------------------------------------------------------------------
// A CTA handles `bpc` blocks one after the other (bpc = 1 for big blocks; several for the small source blocks of a row
// strip of a distributed operator, ~150 KB each at 8 GPUs): the producer lane streams their stages back to back through
// the same ring, so the bulk-copy pipeline is filled once per CTA and never drains at a block boundary; only the
// consumers meet there (the x sub-vector is re-staged). CTA c takes order[c], order[c + G], order[c + 2 G], ... (G = grid):
// the order is heaviest first, so every CTA gets a similar mix.

template <typename T, bool CONJ>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 16 ? 3 : 0) reduce_kernel(KernelSide ks, PassArgs<T> a, int n_blocks, int bpc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SmemLayout sm = carve(smem_raw, ks, ks.stage_bytes, sizeof(T) * ks.block_rows);
    T *xin              = reinterpret_cast<T *>(sm.vec);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    init_barriers(ks, sm);
    pdl_launch_dependents();
    pdl_wait(); // everything below reads or writes vectors / scratch shared with the previous kernel of the stream
    __syncthreads();
    // the producer starts streaming coefficients at once: the first bulk copies are in flight while the consumers
    // wait for / stage the block's x sub-vector (a consumer-only named barrier orders that hand-off)
    if (warp == kConsumerWarps) {
        if (lane == 0) {
            const uint64_t policy = ks.evict_first ? l2_evict_first_policy() : 0;
            RingPos pos;
            for (int bi = 0; bi < bpc; bi++) {
                const int slot_id = static_cast<int>(blockIdx.x + bi * gridDim.x);
                if (slot_id >= n_blocks)
                    break;
                const BlockDesc bd = ks.blocks[ks.order[slot_id]];
                if (bd.n_stages == 0)
                    continue;
                StageDesc next = ks.stages[bd.first_stage];
                for (uint32_t st = 0; st < bd.n_stages; st++) {
                    const StageDesc sd = next;
                    if (st + 1 < bd.n_stages)
                        next = ks.stages[bd.first_stage + st + 1]; // in flight while this stage waits for its slot
                    if (a.twice_only && !(sd.flags & 1u))
                        continue;
                    const uint32_t full = smem_u32(&sm.full[pos.slot]);
                    mbar_wait(smem_u32(&sm.empty[pos.slot]), pos.phase ^ 1u);
                    mbar_arrive_expect_tx(full, sd.nbytes);
                    bulk_g2s(smem_u32(sm.ring + static_cast<size_t>(pos.slot) * sm.slot_bytes), ks.stream + sd.byte_off, sd.nbytes, full, policy, ks.evict_first != 0);
                    pos.advance(ks.ring_stages);
                }
            }
        }
        return;
    }
    RingPos pos;
    uint32_t ubase = warp;
    for (int bi = 0; bi < bpc; bi++) {
        const int slot_id = static_cast<int>(blockIdx.x + bi * gridDim.x);
        if (slot_id >= n_blocks)
            break;
        const uint32_t block_id    = ks.order[slot_id];
        const BlockDesc bd         = ks.blocks[block_id];
        const uint32_t n_my_stages = a.twice_only ? bd.n_twice_stages : bd.n_stages;
        if (n_my_stages == 0)
            continue;
        if (bi > 0)
            consumer_barrier(); // every consumer has finished the units of the previous block: xin may be overwritten
        if (a.wait_flags) { // distributed: the slice of x this block reads is written by its owner's push kernel (dist.cu)
            if (threadIdx.x == 0) {
                const uint32_t ow = a.wait_owner[block_id];
                if (ow != 0xFFFFFFFFu)
                    for (uint32_t q = ow & 0xFFFFu; q <= (ow >> 16); q++)
                        while (ld_acquire_sys(a.wait_flags + q) < a.wait_epoch)
                            __nanosleep(64);
            }
            consumer_barrier();
        }
        // stage the block's x sub-vector in shared memory (every unit of the block multiplies a slice of it)
        for (int i = threadIdx.x; i < ks.block_rows; i += kConsumerWarps * 32) {
            const long long g = static_cast<long long>(bd.row_start) + i + a.in_shift;
            xin[i]            = (i < bd.nrows && g >= 0 && g < a.in_len) ? a.in[g * a.stride] : zero_of(T{});
        }
        consumer_barrier();

        // Every warp walks every stage (in the producer's order); the units are dealt round-robin over the warps ACROSS
        // stages (ubase), so that a stage with few units does not always land on the same warps.
        for (uint32_t q = 0; q < n_my_stages; q++, pos.advance(ks.ring_stages)) {
            const uint32_t slot = pos.slot;
            mbar_wait(smem_u32(&sm.full[slot]), pos.phase);
            const unsigned char *stage = sm.ring + static_cast<size_t>(slot) * sm.slot_bytes;
            const StageHeader hdr      = *reinterpret_cast<const StageHeader *>(stage);
            const Unit *units          = reinterpret_cast<const Unit *>(stage + sizeof(StageHeader));
            const T *data              = reinterpret_cast<const T *>(stage + hdr.data_byte_off);
            // ADDVEC units (dense leaves, direction 0): hand their x slices to the APPLY pass through the c-stream. They
            // hold no coefficients: one unit per LANE, over all the consumer lanes of the CTA.
            for (uint32_t u = hdr.n_panel + warp * 32 + lane; u < hdr.n_units; u += kConsumerWarps * 32) {
                const Unit un = units[u];
                if (a.twice_only && !unit_twice(un.geom))
                    continue;
                const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom);
                T *out = a.scratch + un.out;
                for (uint32_t i = 0; i < h; i++)
                    out[i] = xin[row0 + i];
            }
            for (uint32_t u = ubase; u < hdr.n_panel; u += kConsumerWarps) {
                const Unit un = units[u];
                if (a.twice_only && !unit_twice(un.geom))
                    continue;
                reduce_unit<T, CONJ>(un, data, xin, a.scratch, lane);
            }
            ubase = (ubase - hdr.n_panel) & (kConsumerWarps - 1); // == (warp - units dealt so far) mod 8
            __syncwarp();
            if (lane == 0)
                mbar_arrive(smem_u32(&sm.empty[slot]));
        }
    }
}

// ---- APPLY --------------------------------------------------------------------------------------------
// FUSED (symmetric / Hermitian storage, north_star item 4): while a stage is in shared memory for y += op(P) c, the units
// of the leaves stored once are ALSO reduced against the second input (t' = op2(P)^T x2): the transposed second
// application reads the side's coefficients from the same bulk copy instead of streaming them again.
template <typename T, bool CONJ, bool FUSED, bool CONJ2>
__global__ void __launch_bounds__(kThreads, FUSED ? (sizeof(T) == 16 ? 2 : 3) : 0) apply_kernel(KernelSide ks, PassArgs<T> a) {
    constexpr int R = Rows<T>::R;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const BlockDesc bd = ks.blocks[ks.order[blockIdx.x]];
    if (a.twice_only && !(bd.flags & 1u))
        return; // accumulate-only pass and nothing to add
    const uint32_t n_my_stages = a.twice_only ? bd.n_twice_stages : bd.n_stages;
    const SmemLayout sm = carve(smem_raw, ks, ks.stage_bytes + ks.cseg_bytes, sizeof(T) * ks.block_rows * (kConsumerWarps + (FUSED ? 1 : 0)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *yacc_all = reinterpret_cast<T *>(sm.vec);
    T *xin      = yacc_all + static_cast<size_t>(ks.block_rows) * kConsumerWarps; // FUSED only

    init_barriers(ks, sm);
    pdl_launch_dependents();
    pdl_wait(); // the c-stream and the partials were written by the previous kernels of the stream
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (lane == 0)
            produce<T, true>(ks, bd, sm, a.twice_only, a.scratch + ks.cs_base);
    } else {
        // (the producer is already streaming) every warp clears its own accumulator; FUSED: the consumers stage x2
        T *yacc = yacc_all + static_cast<size_t>(warp) * ks.block_rows;
        for (int i = lane; i < ks.block_rows; i += 32)
            yacc[i] = zero_of(T{});
        if (FUSED) {
            for (int i = threadIdx.x; i < ks.block_rows; i += kConsumerWarps * 32) {
                const long long g = static_cast<long long>(bd.row_start) + i + a.in_shift;
                xin[i]            = (i < bd.nrows && g >= 0 && g < a.in_len) ? a.in[g * a.stride] : zero_of(T{});
            }
            consumer_barrier();
        }
        __syncwarp();
        // Every warp walks every stage; units are dealt round-robin over the warps across stages. The deal is a
        // function of the stream only, so the split of the y contributions over the warps (hence the rounding) is
        // the same in every run.
        RingPos pos;
        uint32_t ubase = warp;
        for (uint32_t q = 0; q < n_my_stages; q++, pos.advance(ks.ring_stages)) {
            const uint32_t slot = pos.slot;
            mbar_wait(smem_u32(&sm.full[slot]), pos.phase);
            const unsigned char *stage = sm.ring + static_cast<size_t>(slot) * sm.slot_bytes;
            const StageHeader hdr      = *reinterpret_cast<const StageHeader *>(stage);
            const Unit *units          = reinterpret_cast<const Unit *>(stage + sizeof(StageHeader));
            const T *data              = reinterpret_cast<const T *>(stage + hdr.data_byte_off);
            const T *cseg              = reinterpret_cast<const T *>(stage + ks.stage_bytes);
            // ADDVEC units: dense leaves applied transposed, their z = op(A)^T x was produced by the REDUCE pass of side 0
            for (uint32_t u = hdr.n_panel + warp; u < hdr.n_units; u += kConsumerWarps) {
                const Unit un = units[u];
                if (a.twice_only && !unit_twice(un.geom))
                    continue;
                const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom);
                const T *c = cseg + un.cslot;
                for (uint32_t i = lane; i < h; i += 32)
                    yacc[row0 + i] = add(yacc[row0 + i], c[i]);
                if (FUSED && unit_twice(un.geom)) { // second application, direction 0: hand the x2 slice over
                    T *out = a.scratch2 + un.out;
                    for (uint32_t i = lane; i < h; i += 32)
                        out[i] = xin[row0 + i];
                }
            }
            for (uint32_t u = ubase; u < hdr.n_panel; u += kConsumerWarps) {
                const Unit un = units[u];
                if (a.twice_only && !unit_twice(un.geom))
                    continue;
                if (FUSED && unit_twice(un.geom)) { // both applications from one walk over the panel
                    fused_unit<T, CONJ, CONJ2>(un, data, cseg + un.cslot, xin, yacc, a.scratch2, lane);
                    continue;
                }
                const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom), w = unit_w(un.geom);
                const T *c = cseg + un.cslot;
                const T *P        = data + un.data_off;
                const uint32_t ld = unit_ld(h, sizeof(T));
                const LaneMap m   = lane_map<R>(h, lane);
                T acc[kMaxQ][R];
#pragma unroll
                for (int qq = 0; qq < kMaxQ; qq++)
#pragma unroll
                    for (int r = 0; r < R; r++)
                        acc[qq][r] = zero_of(T{});
                const uint32_t i0   = R * m.li;
                const uint32_t off0 = i0 < ld ? i0 : 0u;
                if (m.logG == 0) {
                    // tall unit: lanes along the rows, every lane walks all the columns
                    const uint32_t i1   = i0 + 32u * R;
                    const uint32_t off1 = i1 < ld ? i1 : 0u;
                    if (m.Q == 1) {
#pragma unroll 4
                        for (uint32_t k = 0; k < w; k++) {
                            T f[R];
                            load_rows(P + k * ld + off0, f);
                            const T ck = c[k];
#pragma unroll
                            for (int r = 0; r < R; r++)
                                acc[0][r] = fma_(cj<CONJ>(f[r]), ck, acc[0][r]);
                        }
                    } else {
#pragma unroll 4
                        for (uint32_t k = 0; k < w; k++) {
                            T f[R], f1[R];
                            load_rows(P + k * ld + off0, f);
                            load_rows(P + k * ld + off1, f1);
                            const T ck = c[k];
#pragma unroll
                            for (int r = 0; r < R; r++) {
                                acc[0][r] = fma_(cj<CONJ>(f[r]), ck, acc[0][r]);
                                acc[1][r] = fma_(cj<CONJ>(f1[r]), ck, acc[1][r]);
                            }
                        }
                    }
                } else {
                    // short unit: the G segments of lanes work on interleaved columns, then a butterfly over the segments
                    for (uint32_t kk = 0; kk < w; kk += 1u << m.logG) {
                        const uint32_t k  = kk + m.g;
                        const bool valid  = k < w;
                        const uint32_t kc = valid ? k : 0u;
                        T f[R];
                        load_rows(P + kc * ld + off0, f);
                        const T ck = valid ? c[kc] : zero_of(T{});
#pragma unroll
                        for (int r = 0; r < R; r++)
                            acc[0][r] = fma_(cj<CONJ>(f[r]), ck, acc[0][r]);
                    }
                    for (int d = 1 << m.seglog; d < 32; d <<= 1)
#pragma unroll
                        for (int r = 0; r < R; r++)
                            acc[0][r] = add(acc[0][r], shfl_xor(acc[0][r], d));
                }
                if (m.g == 0) {
#pragma unroll
                    for (int qq = 0; qq < kMaxQ; qq++)
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const uint32_t i = i0 + 32u * R * qq + r;
                            if (i < h && (qq == 0 || m.Q > 1))
                                yacc[row0 + i] = add(yacc[row0 + i], acc[qq][r]);
                        }
                }
            }
            ubase = (ubase - hdr.n_panel) & (kConsumerWarps - 1);
            __syncwarp();
            if (lane == 0)
                mbar_arrive(smem_u32(&sm.empty[slot]));
        }
    }
    __syncthreads();
    // epilogue: fixed-order sum over the warps' private accumulators, alpha / beta, one write per row
    for (int i = threadIdx.x; i < bd.nrows; i += kThreads) {
        T s = zero_of(T{});
#pragma unroll
        for (int wv = 0; wv < kConsumerWarps; wv++)
            s = add(s, yacc_all[static_cast<size_t>(wv) * ks.block_rows + i]);
        const long long g = static_cast<long long>(bd.row_start) + i + a.out_shift;
        if (g >= 0 && g < a.out_len) {
            T r = mul(a.alpha, s);
            if (!a.beta_is_zero)
                r = fma_(a.beta, a.out[g * a.stride], r);
            a.out[g * a.stride] = r;
        }
    }
}

// ---- small kernels ------------------------------------------------------------------------------------
// One warp per piece. Phase 1: v[k] = sum of the piece's partials; the 32 / len2 lane groups (len2 = len rounded up to a
// power of two) take the chunks round-robin and a butterfly folds the groups: a fixed order that depends only on
// (len, n_sum), hence deterministic. Phase 2: v (or the sub-range a consumer asked for) is written into every
// consumer slot of the c-stream; pieces with many consumers (large leaves span hundreds of blocks) spread them over
// the lanes instead of walking them one by one.
template <typename T>
__global__ void combine_kernel(const CombineEntry *entries, const CombineDst *dsts, int n, T *scratch, int twice_only) {
    const int warps_per_block = blockDim.x >> 5;
    const int e               = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    pdl_launch_dependents();
    pdl_wait();
    if (e >= n)
        return;
    const CombineEntry ce = entries[e];
    if (twice_only && !combine_twice(ce.packed))
        return;
    const uint32_t lane = threadIdx.x & 31, len = combine_len(ce.packed), n_sum = combine_n_sum(ce.packed);
    const int log2len   = len > 16 ? 5 : (len > 8 ? 4 : (len > 4 ? 3 : (len > 2 ? 2 : (len > 1 ? 1 : 0))));
    const uint32_t k = lane & ((1u << log2len) - 1u), jg = lane >> log2len, G = 32u >> log2len;
    T v = zero_of(T{});
    if (k < len) {
        const T *p = scratch + ce.src + k;
        // (deep unroll: the loads are independent, the adds keep their order; a piece of a 32k-row leaf sums 256 partials
        // and the slowest warp sets the duration of this latency-bound kernel)
#pragma unroll 16
        for (uint32_t j = jg; j < n_sum; j += G)
            v = add(v, p[static_cast<size_t>(j) * len]);
    }
    for (int d = 1 << log2len; d < 32; d <<= 1)
        v = add(v, shfl_xor(v, d));
    // every lane now holds v[lane % len2]
    if (ce.n_dst <= 4) {
        for (uint32_t q = 0; q < ce.n_dst; q++) {
            const CombineDst d = dsts[ce.dst_first + q];
            if (lane >= d.sub_off && lane < static_cast<uint32_t>(d.sub_off) + d.sub_len)
                scratch[d.slot + lane - d.sub_off] = v;
        }
    } else {
        for (uint32_t q0 = 0; q0 < ce.n_dst; q0 += 32) {
            const uint32_t q   = q0 + lane;
            const bool mine    = q < ce.n_dst;
            const CombineDst d = mine ? dsts[ce.dst_first + q] : CombineDst{0u, 0, 0};
            for (uint32_t el = 0; el < len; el++) {
                const T val = shfl_xor(v, static_cast<int>(lane ^ el)); // value of lane el
                if (mine && el >= d.sub_off && el < static_cast<uint32_t>(d.sub_off) + d.sub_len)
                    scratch[d.slot + el - d.sub_off] = val;
            }
        }
    }
}

template <typename T>
__global__ void permute_kernel(const T *in, T *out, const int32_t *perm, int n, int mu, int gather, int colmajor_user) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(n) * mu)
        return;
    const int i = static_cast<int>(idx / mu), c = static_cast<int>(idx % mu);
    const int p = perm[i];
    // cluster side is row-major (i*mu + c); user side is row-major or column-major (c*n + i)
    const long long u = colmajor_user ? static_cast<long long>(c) * n + p : static_cast<long long>(p) * mu + c;
    if (gather) // cluster[i] = user[perm[i]]
        out[idx] = in[u];
    else // user[perm[i]] = cluster[i]
        out[u] = in[idx];
}

template <typename T>
__global__ void scale_kernel(T *y, long long n, T beta, int beta_is_zero) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        y[i] = beta_is_zero ? zero_of(T{}) : mul(beta, y[i]);
}

__global__ void wait_flags_kernel(const unsigned long long *flags, int world, unsigned long long epoch) {
    if (static_cast<int>(threadIdx.x) < world)
        while (ld_acquire_sys(flags + threadIdx.x) < epoch)
            __nanosleep(64);
}

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream drains its last
// wave; its CTAs set up their barriers, then block in griddepcontrol.wait until the predecessor has completed and
// flushed (pdl_wait() below), so the launch latency and the CTA ramp-up at every pass boundary are hidden.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3(static_cast<unsigned>(grid));
    cfg.blockDim           = dim3(static_cast<unsigned>(block));
    cfg.dynamicSmemBytes   = smem;
    cfg.stream             = st;
    cudaLaunchAttribute attr;
    attr.id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
    cfg.attrs                                       = &attr;
    cfg.numAttrs                                    = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline KernelSide make_kernel_side(const SideDevice &s, const LaunchConfig &cfg, int ring) {
    return KernelSide{s.blocks, s.stages, s.order, s.stream, s.cs_base, cfg.block_rows, cfg.stage_bytes, cfg.cseg_bytes, ring, cfg.evict_first};
}

inline bool is_zero(double v) { return v == 0.; }
inline bool is_zero(cplx v) { return v.x == 0. && v.y == 0.; }

template <typename T>
struct Kernels;
template <>
struct Kernels<double> {
    static void reduce(const KernelSide &ks, const PassArgs<double> &a, int n_blocks, int bpc, size_t smem, cudaStream_t st) {
        launch_pdl(reduce_kernel<double, false>, (n_blocks + bpc - 1) / bpc, kThreads, smem, st, ks, a, n_blocks, bpc);
    }
    static void apply(const KernelSide &ks, const PassArgs<double> &a, int grid, size_t smem, cudaStream_t st) {
        if (a.fused)
            launch_pdl(apply_kernel<double, false, true, false>, grid, kThreads, smem, st, ks, a);
        else
            launch_pdl(apply_kernel<double, false, false, false>, grid, kThreads, smem, st, ks, a);
    }
    static cudaError_t configure(int rs, int as, int fs) {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(reduce_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs)) != cudaSuccess)
            return e;
        if ((e = cudaFuncSetAttribute(apply_kernel<double, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, as)) != cudaSuccess)
            return e;
        return cudaFuncSetAttribute(apply_kernel<double, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs);
    }
};
template <>
struct Kernels<cplx> {
    static void reduce(const KernelSide &ks, const PassArgs<cplx> &a, int n_blocks, int bpc, size_t smem, cudaStream_t st) {
        const int grid = (n_blocks + bpc - 1) / bpc;
        if (a.conj)
            launch_pdl(reduce_kernel<cplx, true>, grid, kThreads, smem, st, ks, a, n_blocks, bpc);
        else
            launch_pdl(reduce_kernel<cplx, false>, grid, kThreads, smem, st, ks, a, n_blocks, bpc);
    }
    static void apply(const KernelSide &ks, const PassArgs<cplx> &a, int grid, size_t smem, cudaStream_t st) {
        if (a.fused) { // (conj, conj2): (0,0) symmetric, (0,1) Hermitian 'N', (1,0) Hermitian 'C'
            if (a.conj)
                launch_pdl(apply_kernel<cplx, true, true, false>, grid, kThreads, smem, st, ks, a);
            else if (a.conj2)
                launch_pdl(apply_kernel<cplx, false, true, true>, grid, kThreads, smem, st, ks, a);
            else
                launch_pdl(apply_kernel<cplx, false, true, false>, grid, kThreads, smem, st, ks, a);
        } else if (a.conj)
            launch_pdl(apply_kernel<cplx, true, false, false>, grid, kThreads, smem, st, ks, a);
        else
            launch_pdl(apply_kernel<cplx, false, false, false>, grid, kThreads, smem, st, ks, a);
    }
    static cudaError_t configure(int rs, int as, int fs) {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(reduce_kernel<cplx, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs)) != cudaSuccess)
            return e;
        if ((e = cudaFuncSetAttribute(reduce_kernel<cplx, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs)) != cudaSuccess)
            return e;
        if ((e = cudaFuncSetAttribute(apply_kernel<cplx, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, as)) != cudaSuccess)
            return e;
        if ((e = cudaFuncSetAttribute(apply_kernel<cplx, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, as)) != cudaSuccess)
            return e;
        if ((e = cudaFuncSetAttribute(apply_kernel<cplx, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs)) != cudaSuccess)
            return e;
        if ((e = cudaFuncSetAttribute(apply_kernel<cplx, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs)) != cudaSuccess)
            return e;
        return cudaFuncSetAttribute(apply_kernel<cplx, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs);
    }
};

} // namespace

size_t reduce_smem_bytes(const LaunchConfig &cfg, size_t esize) {
    const size_t slots = static_cast<size_t>(cfg.reduce_ring_stages);
    return slots * cfg.stage_bytes + esize * cfg.block_rows + 16 * slots;
}
size_t apply_smem_bytes(const LaunchConfig &cfg, size_t esize) {
    const size_t slots = static_cast<size_t>(cfg.ring_stages);
    return slots * (cfg.stage_bytes + cfg.cseg_bytes) + esize * cfg.block_rows * kConsumerWarps + 16 * slots;
}

// The fused complex kernel is register-limited to 2 CTAs per SM (the others run 3): the shared memory that frees
// pays for one more ring slot (measured: Helmholtz N = 1e6 'S', 7.78 -> 7.13 ms)
static int fused_ring_stages(const LaunchConfig &cfg, size_t esize) { return cfg.ring_stages + (esize == 16 ? 1 : 0); }
size_t fused_smem_bytes(const LaunchConfig &cfg, size_t esize) {
    const size_t slots = static_cast<size_t>(fused_ring_stages(cfg, esize));
    return slots * (cfg.stage_bytes + cfg.cseg_bytes) + esize * cfg.block_rows * (kConsumerWarps + 1) + 16 * slots;
}

cudaError_t configure_kernels(const LaunchConfig &cfg) {
    cudaError_t e = Kernels<double>::configure(static_cast<int>(reduce_smem_bytes(cfg, 8)), static_cast<int>(apply_smem_bytes(cfg, 8)), static_cast<int>(fused_smem_bytes(cfg, 8)));
    if (e != cudaSuccess)
        return e;
    return Kernels<cplx>::configure(static_cast<int>(reduce_smem_bytes(cfg, 16)), static_cast<int>(apply_smem_bytes(cfg, 16)), static_cast<int>(fused_smem_bytes(cfg, 16)));
}

template <typename T>
cudaError_t launch_reduce(const SideDevice &side, const LaunchConfig &cfg, const PassArgs<T> &args, cudaStream_t stream) {
    if (side.n_blocks == 0)
        return cudaSuccess;
    // blocks per CTA: several when the side's blocks are small (a row strip of a distributed operator), else 1; option
    // reduce_blocks_per_cta. Measured on a 1/8 strip of N = 1e6 (142 KB per block): 1 -> 0.220 ms, 2 -> 0.208, 4 -> 0.204.
    int bpc = cfg.reduce_blocks_per_cta;
    if (bpc <= 0) {
        const uint64_t avg = side.n_blocks > 0 ? side.stream_bytes / static_cast<uint64_t>(side.n_blocks) : 0;
        bpc                = side.n_blocks < 2 ? 1 : (avg < (uint64_t(256) << 10) ? 4 : (avg < (uint64_t(800) << 10) ? 2 : 1));
    }
    if (bpc > 8)
        bpc = 8;
    Kernels<T>::reduce(make_kernel_side(side, cfg, cfg.reduce_ring_stages), args, side.n_blocks, bpc, reduce_smem_bytes(cfg, sizeof(T)), stream);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_apply(const SideDevice &side, const LaunchConfig &cfg, const PassArgs<T> &args, cudaStream_t stream) {
    if (side.n_blocks == 0)
        return cudaSuccess;
    PassArgs<T> a  = args;
    a.beta_is_zero = is_zero(args.beta) ? 1 : 0;
    if (a.fused)
        Kernels<T>::apply(make_kernel_side(side, cfg, fused_ring_stages(cfg, sizeof(T))), a, side.n_blocks, fused_smem_bytes(cfg, sizeof(T)), stream);
    else
        Kernels<T>::apply(make_kernel_side(side, cfg, cfg.ring_stages), a, side.n_blocks, apply_smem_bytes(cfg, sizeof(T)), stream);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_combine(const SideDevice &side, T *scratch, int twice_only, cudaStream_t stream) {
    if (side.n_combine == 0)
        return cudaSuccess;
    const int warps = 8;
    launch_pdl(combine_kernel<T>, (side.n_combine + warps - 1) / warps, warps * 32, 0, stream, side.combine, side.combine_dst, side.n_combine, scratch, twice_only);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_permute(const T *in, T *out, const int32_t *perm, int n, int mu, bool gather, bool colmajor_user, cudaStream_t stream) {
    const long long total = static_cast<long long>(n) * mu;
    if (total == 0)
        return cudaSuccess;
    permute_kernel<T><<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, out, perm, n, mu, gather ? 1 : 0, colmajor_user ? 1 : 0);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_scale(T *y, long long n, T beta, cudaStream_t stream) {
    if (n == 0)
        return cudaSuccess;
    scale_kernel<T><<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(y, n, beta, is_zero(beta) ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_wait_flags(const unsigned long long *flags, int world, unsigned long long epoch, cudaStream_t stream) {
    wait_flags_kernel<<<1, 32, 0, stream>>>(flags, world, epoch);
    return cudaGetLastError();
}

#define HTB_INSTANTIATE(T)                                                                                              \
    template cudaError_t launch_reduce<T>(const SideDevice &, const LaunchConfig &, const PassArgs<T> &, cudaStream_t); \
    template cudaError_t launch_apply<T>(const SideDevice &, const LaunchConfig &, const PassArgs<T> &, cudaStream_t);  \
    template cudaError_t launch_combine<T>(const SideDevice &, T *, int, cudaStream_t);                                 \
    template cudaError_t launch_permute<T>(const T *, T *, const int32_t *, int, int, bool, bool, cudaStream_t);        \
    template cudaError_t launch_scale<T>(T *, long long, T, cudaStream_t);
HTB_INSTANTIATE(double)
HTB_INSTANTIATE(cplx)

} // namespace htb
