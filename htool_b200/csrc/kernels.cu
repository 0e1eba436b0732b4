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
// UBLKCP) that a single producer lane issues and mbarriers track. Eight consumer warps walk the units
// of each stage out of shared memory:
//   REDUCE  t[k] = sum_i op(P[i,k]) x[i]: the block's x sub-vector is staged in shared memory once,
//           per-lane FMAs then a fixed-order warp-shuffle butterfly;
//   APPLY   y[i] += sum_k op(P[i,k]) c[k]: per-warp private y accumulators in shared memory, summed over
//           warps in warp order at the end and written once (y rows are owned by exactly one CTA):
//           deterministic, no atomics, beta/alpha fused in the same epilogue.
// Bandwidth-bound: every coefficient crosses HBM->SMEM once per pass and is used for one FMA.
#include "kernels.cuh"

#include <cstdint>

namespace htb {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kThreads       = (kConsumerWarps + 1) * 32; // + 1 producer warp
constexpr int kMaxQ          = 4;                         // block_rows <= 128 -> <= 4 rows per lane

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
__device__ __forceinline__ double cj(double a, int) { return a; }
__device__ __forceinline__ cplx cj(cplx a, int conj) { return conj ? cplx{a.x, -a.y} : a; }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ cplx shfl_xor(cplx v, int m) { return cplx{__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m)}; }
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1)
        v = add(v, shfl_xor(v, m));
    return v;
}

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
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

struct KernelSide {
    const BlockDesc *blocks;
    const StageDesc *stages;
    const uint32_t *order;
    const unsigned char *stream;
    int block_rows, stage_bytes, ring_stages, evict_first;
};

// Shared-memory carve-up common to both kernels: [ring | vec | cbuf | barriers]
struct SmemLayout {
    unsigned char *ring;
    unsigned char *vec;  // REDUCE: x sub-vector (block_rows). APPLY: per-warp y accumulators (warps x block_rows)
    unsigned char *cbuf; // APPLY: per-warp c vector (warps x 32)
    uint64_t *full, *empty;
};
__device__ __forceinline__ SmemLayout carve(unsigned char *base, const KernelSide &ks, size_t vec_bytes, size_t cbuf_bytes) {
    SmemLayout s;
    s.ring  = base;
    s.vec   = base + static_cast<size_t>(ks.ring_stages) * ks.stage_bytes;
    s.cbuf  = s.vec + vec_bytes;
    s.full  = reinterpret_cast<uint64_t *>(s.cbuf + cbuf_bytes);
    s.empty = s.full + ks.ring_stages;
    return s;
}

// Producer: one lane streams the block's stages into the ring.
__device__ __forceinline__ void produce(const KernelSide &ks, const BlockDesc &bd, const SmemLayout &sm, int twice_only) {
    const uint64_t policy = ks.evict_first ? l2_evict_first_policy() : 0;
    uint32_t it           = 0;
    for (uint32_t st = 0; st < bd.n_stages; st++) {
        const StageDesc sd = ks.stages[bd.first_stage + st];
        if (twice_only && !(sd.flags & 1u))
            continue;
        const uint32_t slot = it % ks.ring_stages, round = it / ks.ring_stages;
        mbar_wait(smem_u32(&sm.empty[slot]), (round & 1u) ^ 1u);
        mbar_arrive_expect_tx(smem_u32(&sm.full[slot]), sd.nbytes);
        bulk_g2s(smem_u32(sm.ring + static_cast<size_t>(slot) * ks.stage_bytes), ks.stream + sd.byte_off, sd.nbytes, smem_u32(&sm.full[slot]), policy, ks.evict_first != 0);
        it++;
    }
}

__device__ __forceinline__ void init_barriers(const KernelSide &ks, const SmemLayout &sm) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < ks.ring_stages; s++) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

// ---- REDUCE -------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) reduce_kernel(KernelSide ks, PassArgs<T> a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const BlockDesc bd = ks.blocks[ks.order[blockIdx.x]];
    if (bd.n_stages == 0 || (a.twice_only && !(bd.flags & 1u)))
        return;
    const SmemLayout sm = carve(smem_raw, ks, sizeof(T) * ks.block_rows, 0);
    T *xin              = reinterpret_cast<T *>(sm.vec);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    init_barriers(ks, sm);
    // stage the block's x sub-vector in shared memory (every unit of the block multiplies a slice of it)
    for (int i = threadIdx.x; i < ks.block_rows; i += kThreads) {
        const long long g = static_cast<long long>(bd.row_start) + i + a.in_shift;
        xin[i]            = (i < bd.nrows && g >= 0 && g < a.in_len) ? a.in[g * a.stride] : zero_of(T{});
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (lane == 0)
            produce(ks, bd, sm, a.twice_only);
        return;
    }

    uint32_t it = 0;
    for (uint32_t st = 0; st < bd.n_stages; st++) {
        if (a.twice_only && !(ks.stages[bd.first_stage + st].flags & 1u))
            continue;
        const uint32_t slot = it % ks.ring_stages, round = it / ks.ring_stages;
        mbar_wait(smem_u32(&sm.full[slot]), round & 1u);
        const unsigned char *stage = sm.ring + static_cast<size_t>(slot) * ks.stage_bytes;
        const StageHeader hdr      = *reinterpret_cast<const StageHeader *>(stage);
        const Unit *units          = reinterpret_cast<const Unit *>(stage + sizeof(StageHeader));
        const T *data              = reinterpret_cast<const T *>(stage + hdr.data_byte_off);
        for (uint32_t u = warp; u < hdr.n_units; u += kConsumerWarps) {
            const Unit un       = units[u];
            const uint32_t kind = unit_kind(un.geom);
            if (kind == UNIT_ADDVEC || (a.twice_only && !unit_twice(un.geom)))
                continue;
            const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom), w = unit_w(un.geom);
            const T *P = data + un.data_off;
            T xv[kMaxQ];
#pragma unroll
            for (int q = 0; q < kMaxQ; q++) {
                const uint32_t i = lane + 32u * q;
                xv[q]            = i < h ? xin[row0 + i] : zero_of(T{});
            }
            T mine = zero_of(T{});
            for (uint32_t k = 0; k < w; k++) {
                const T *col = P + static_cast<size_t>(k) * h;
                T s          = zero_of(T{});
#pragma unroll
                for (int q = 0; q < kMaxQ; q++) {
                    const uint32_t i = lane + 32u * q;
                    if (i < h)
                        s = fma_(cj(col[i], a.conj), xv[q], s);
                }
                s = warp_sum(s);
                if (lane == k)
                    mine = s;
            }
            if (lane < w)
                a.scratch[static_cast<size_t>(un.aux_reduce) + lane] = mine;
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(smem_u32(&sm.empty[slot]));
        it++;
    }
}

// ---- APPLY --------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) apply_kernel(KernelSide ks, PassArgs<T> a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const BlockDesc bd = ks.blocks[ks.order[blockIdx.x]];
    if (a.twice_only && !(bd.flags & 1u))
        return; // accumulate-only pass and nothing to add
    const SmemLayout sm = carve(smem_raw, ks, sizeof(T) * ks.block_rows * kConsumerWarps, sizeof(T) * 32 * kConsumerWarps);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *yacc_all = reinterpret_cast<T *>(sm.vec);

    init_barriers(ks, sm);
    for (int i = threadIdx.x; i < ks.block_rows * kConsumerWarps; i += kThreads)
        yacc_all[i] = zero_of(T{});
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (lane == 0)
            produce(ks, bd, sm, a.twice_only);
    } else {
        T *yacc = yacc_all + static_cast<size_t>(warp) * ks.block_rows;
        T *cbuf = reinterpret_cast<T *>(sm.cbuf) + warp * 32;
        uint32_t it = 0;
        for (uint32_t st = 0; st < bd.n_stages; st++) {
            if (a.twice_only && !(ks.stages[bd.first_stage + st].flags & 1u))
                continue;
            const uint32_t slot = it % ks.ring_stages, round = it / ks.ring_stages;
            mbar_wait(smem_u32(&sm.full[slot]), round & 1u);
            const unsigned char *stage = sm.ring + static_cast<size_t>(slot) * ks.stage_bytes;
            const StageHeader hdr      = *reinterpret_cast<const StageHeader *>(stage);
            const Unit *units          = reinterpret_cast<const Unit *>(stage + sizeof(StageHeader));
            const T *data              = reinterpret_cast<const T *>(stage + hdr.data_byte_off);
            for (uint32_t u = warp; u < hdr.n_units; u += kConsumerWarps) {
                const Unit un = units[u];
                if (a.twice_only && !unit_twice(un.geom))
                    continue;
                const uint32_t kind = unit_kind(un.geom);
                const uint32_t row0 = unit_row0(un.geom), h = unit_h(un.geom), w = unit_w(un.geom);
                if (kind == UNIT_ADDVEC) {
                    // dense leaf applied transposed: its z = op(A)^T x was produced by the REDUCE pass of side 0
                    for (uint32_t i = lane; i < h; i += 32)
                        yacc[row0 + i] = add(yacc[row0 + i], a.scratch[static_cast<size_t>(un.aux_apply) + i]);
                    continue;
                }
                // c vector: t (low rank) from scratch, or the x slice (dense) from the input vector
                T cv = zero_of(T{});
                if (lane < w) {
                    if (kind == UNIT_LOWRANK) {
                        cv = a.scratch[static_cast<size_t>(un.aux_apply) + lane];
                    } else {
                        const long long g = static_cast<long long>(un.aux_apply) + lane + a.in_shift;
                        if (g >= 0 && g < a.in_len)
                            cv = a.in[g * a.stride];
                    }
                }
                __syncwarp();
                cbuf[lane] = cv;
                __syncwarp();
                const T *P = data + un.data_off;
                T acc[kMaxQ];
#pragma unroll
                for (int q = 0; q < kMaxQ; q++)
                    acc[q] = zero_of(T{});
                for (uint32_t k = 0; k < w; k++) {
                    const T c    = cbuf[k];
                    const T *col = P + static_cast<size_t>(k) * h;
#pragma unroll
                    for (int q = 0; q < kMaxQ; q++) {
                        const uint32_t i = lane + 32u * q;
                        if (i < h)
                            acc[q] = fma_(cj(col[i], a.conj), c, acc[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < kMaxQ; q++) {
                    const uint32_t i = lane + 32u * q;
                    if (i < h)
                        yacc[row0 + i] = add(yacc[row0 + i], acc[q]);
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(smem_u32(&sm.empty[slot]));
            it++;
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
template <typename T>
__global__ void combine_kernel(const CombineEntry *entries, int n, T *scratch, int twice_only) {
    const int warps_per_block = blockDim.x >> 5;
    const int e               = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    if (e >= n)
        return;
    const CombineEntry ce = entries[e];
    if (twice_only && !(ce.n_chunks & 0x80000000u))
        return;
    const uint32_t nc = ce.n_chunks & 0x7fffffffu;
    for (uint32_t k = threadIdx.x & 31; k < ce.w; k += 32) {
        T s = zero_of(T{});
        for (uint32_t c = 0; c < nc; c++) // chunk order: fixed summation order
            s = add(s, scratch[static_cast<size_t>(ce.src) + static_cast<size_t>(c) * ce.w + k]);
        scratch[static_cast<size_t>(ce.dst) + k] = s;
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
    if (gather) { // cluster[i] = user[perm[i]]
        const long long u = colmajor_user ? static_cast<long long>(c) * n + p : static_cast<long long>(p) * mu + c;
        out[idx]          = in[u];
    } else { // user[perm[i]] = cluster[i]
        const long long u = colmajor_user ? static_cast<long long>(c) * n + p : static_cast<long long>(p) * mu + c;
        out[u]            = in[idx];
    }
}

template <typename T>
__global__ void scale_kernel(T *y, long long n, T beta, int beta_is_zero) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        y[i] = beta_is_zero ? zero_of(T{}) : mul(beta, y[i]);
}

inline KernelSide make_kernel_side(const SideDevice &s, const LaunchConfig &cfg) {
    return KernelSide{s.blocks, s.stages, s.order, s.stream, cfg.block_rows, cfg.stage_bytes, cfg.ring_stages, cfg.evict_first};
}

inline bool is_zero(double v) { return v == 0.; }
inline bool is_zero(cplx v) { return v.x == 0. && v.y == 0.; }

} // namespace

size_t reduce_smem_bytes(const LaunchConfig &cfg, size_t esize) {
    return static_cast<size_t>(cfg.ring_stages) * cfg.stage_bytes + esize * cfg.block_rows + 16 * static_cast<size_t>(cfg.ring_stages);
}
size_t apply_smem_bytes(const LaunchConfig &cfg, size_t esize) {
    return static_cast<size_t>(cfg.ring_stages) * cfg.stage_bytes + esize * cfg.block_rows * kConsumerWarps + esize * 32 * kConsumerWarps + 16 * static_cast<size_t>(cfg.ring_stages);
}

cudaError_t configure_kernels(const LaunchConfig &cfg) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(reduce_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(reduce_smem_bytes(cfg, 8)))) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(reduce_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(reduce_smem_bytes(cfg, 16)))) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(apply_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(apply_smem_bytes(cfg, 8)))) != cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(apply_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(apply_smem_bytes(cfg, 16)))) != cudaSuccess)
        return e;
    return cudaSuccess;
}

template <typename T>
cudaError_t launch_reduce(const SideDevice &side, const LaunchConfig &cfg, const PassArgs<T> &args, cudaStream_t stream) {
    if (side.n_blocks == 0)
        return cudaSuccess;
    reduce_kernel<T><<<side.n_blocks, kThreads, reduce_smem_bytes(cfg, sizeof(T)), stream>>>(make_kernel_side(side, cfg), args);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_apply(const SideDevice &side, const LaunchConfig &cfg, const PassArgs<T> &args, cudaStream_t stream) {
    if (side.n_blocks == 0)
        return cudaSuccess;
    PassArgs<T> a  = args;
    a.beta_is_zero = is_zero(args.beta) ? 1 : 0;
    apply_kernel<T><<<side.n_blocks, kThreads, apply_smem_bytes(cfg, sizeof(T)), stream>>>(make_kernel_side(side, cfg), a);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_combine(const SideDevice &side, T *scratch, int twice_only, cudaStream_t stream) {
    if (side.n_combine == 0)
        return cudaSuccess;
    const int warps = 8;
    combine_kernel<T><<<(side.n_combine + warps - 1) / warps, warps * 32, 0, stream>>>(side.combine, side.n_combine, scratch, twice_only);
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

#define HTB_INSTANTIATE(T)                                                                                                    \
    template cudaError_t launch_reduce<T>(const SideDevice &, const LaunchConfig &, const PassArgs<T> &, cudaStream_t);       \
    template cudaError_t launch_apply<T>(const SideDevice &, const LaunchConfig &, const PassArgs<T> &, cudaStream_t);        \
    template cudaError_t launch_combine<T>(const SideDevice &, T *, int, cudaStream_t);                                       \
    template cudaError_t launch_permute<T>(const T *, T *, const int32_t *, int, int, bool, bool, cudaStream_t);              \
    template cudaError_t launch_scale<T>(T *, long long, T, cudaStream_t);
HTB_INSTANTIATE(double)
HTB_INSTANTIATE(cplx)

} // namespace htb
