// htool_b200/csrc/kernels.cuh — launch interface of the sm_100a kernels (kernels.cu).
#ifndef HTB_KERNELS_CUH
#define HTB_KERNELS_CUH

#include "store.hpp"
#include <cuda_runtime.h>

namespace htb {

struct __align__(16) cplx {
    double x, y;
};

// Device view of one side of the store.
struct SideDevice {
    const BlockDesc *blocks         = nullptr;
    const StageDesc *stages         = nullptr;
    const uint32_t *order           = nullptr;
    const unsigned char *stream     = nullptr;
    const CombineEntry *combine     = nullptr; // direction whose consumer is this side
    const CombineDst *combine_dst   = nullptr;
    const MUnit *munits             = nullptr; // multi-RHS side table, stage order
    const unsigned char *aux_reduce = nullptr; // multi-RHS aux records (runs + column tables), REDUCE_M role
    const unsigned char *aux_apply  = nullptr; // same offsets, APPLY_M role
    const CombineEntry *combine_m   = nullptr; // multi-RHS partial sums of the direction whose consumer is this side
    int n_combine_m                 = 0;
    int n_blocks                    = 0;
    uint64_t stream_bytes           = 0; // bytes of the side's stream (average block size decides how REDUCE groups blocks)
    int n_combine                   = 0;
    int n                           = 0;
    uint64_t cs_base                = 0; // element offset of CS[side] inside a scratch copy
    bool any_twice                  = false;
};

struct LaunchConfig {
    int block_rows  = 128;
    int stage_bytes = 24576;
    int cseg_bytes  = 4096;
    int ring_stages = 2;        // APPLY ring depth (slot = stage + c segment)
    int reduce_ring_stages = 3; // REDUCE ring depth (slot = stage)
    int m_ring_stages        = 4; // APPLY_M ring depth (slot = stage + aux record)
    int m_reduce_ring_stages = 4; // REDUCE_M ring depth (1 CTA / SM: the X block takes 76 KiB)
    int m_reduce_warps       = 24; // REDUCE_M consumer warps (4 .. 24)
    int m_aux_bytes          = 0; // aux part of a ring slot of the multi-RHS kernels: the largest aux record of the store, rounded up to 128
                                  // (0: the worst case a stage can hold, aux_slot_bytes(cseg_bytes))
    int m_pad                = 4; // vector stride of the multi-RHS scratch / B ring / X block = vs + m_pad (mkernels.cuh)
    int m_x_rows             = 0; // tallest block of the store: rows of REDUCE_M's X block in shared memory
    int m_b_producers        = 3; // APPLY_M: B producer warps (1 .. 3)
    int m_b_ring_log2        = 3; // APPLY_M: the ring of B-row chunks (32 rows each) holds 2^this chunks
    int reduce_blocks_per_cta = 0; // REDUCE: blocks handled by one CTA through one ring (0 = automatic: 2 for small blocks, else 1)
    int evict_first = 1; // L2 evict_first hint on the coefficient stream
};

// One pass over a side. Vectors are addressed as v[index * stride + column] (stride = mu, column = RHS).
template <typename T>
struct PassArgs {
    const T *in      = nullptr; // REDUCE: multiplied vector
    long long in_len = 0;
    int in_shift     = 0; // REDUCE: in index = block index + in_shift
    T *out           = nullptr; // APPLY only
    long long out_len = 0;
    int out_shift    = 0; // APPLY: out index = block index + out_shift
    T *scratch       = nullptr; // one scratch copy: REDUCE writes unit results, APPLY reads its c-stream
    T alpha{}, beta{};
    int beta_is_zero = 0; // APPLY: do not read out
    int twice_only   = 0; // only units of leaves applied twice (second, transposed application of symmetric storage)
    int conj         = 0; // conjugate the coefficients (trans == 'C', Hermitian second application)
    int stride       = 1; // distance between consecutive vector entries (mu for row-major multi-RHS)
    // fused APPLY + REDUCE (symmetric storage): the units of the leaves stored once are also reduced against `in`
    // (with in_shift / in_len above) into scratch2, conjugated when conj2
    int fused        = 0;
    T *scratch2      = nullptr;
    int conj2        = 0;
    // distributed REDUCE over peer memory (dist.cu): before a block stages its x sub-vector it waits until the ranks
    // owning that index range have published their slice of x (arrival flags written over NVLink, >= wait_epoch)
    const unsigned long long *wait_flags = nullptr; // [world], this rank's copy
    const uint32_t *wait_owner           = nullptr; // per block id: first owner | last owner << 16, 0xFFFFFFFF = own partition
    unsigned long long wait_epoch        = 0;
};

template <typename T>
cudaError_t launch_reduce(const SideDevice &side, const LaunchConfig &cfg, const PassArgs<T> &args, cudaStream_t stream);
template <typename T>
cudaError_t launch_apply(const SideDevice &side, const LaunchConfig &cfg, const PassArgs<T> &args, cudaStream_t stream);
// Folds the partials of the direction whose consumer is `side` and replicates them into the consumer slots.
template <typename T>
cudaError_t launch_combine(const SideDevice &side, T *scratch, int twice_only, cudaStream_t stream);

// out[i] = in[perm[i]] (gather) / out[perm[i]] = in[i] (scatter), i < n, for mu interleaved columns:
// cluster_node.hpp:150-175 (user_to_cluster / cluster_to_user) on the device.
// colmajor_user: the non-permuted side is column-major n x mu (user layout of add_hmatrix_matrix_product)
// while the permuted side is row-major (mu contiguous).
template <typename T>
cudaError_t launch_permute(const T *in, T *out, const int32_t *perm, int n, int mu, bool gather, bool colmajor_user, cudaStream_t stream);
// y <- beta * y (used when an operator has no block at all on the output side)
template <typename T>
cudaError_t launch_scale(T *y, long long n, T beta, cudaStream_t stream);

// one warp spinning until flags[q] >= epoch for every q < world (peer-memory gather, multi-RHS path)
cudaError_t launch_wait_flags(const unsigned long long *flags, int world, unsigned long long epoch, cudaStream_t stream);

extern bool g_pdl; // programmatic dependent launch of the pass kernels (option "pdl")

size_t reduce_smem_bytes(const LaunchConfig &cfg, size_t esize);
size_t apply_smem_bytes(const LaunchConfig &cfg, size_t esize);
size_t fused_smem_bytes(const LaunchConfig &cfg, size_t esize);
cudaError_t configure_kernels(const LaunchConfig &cfg); // sets the dynamic shared memory attributes once

} // namespace htb
#endif
