// htool_b200/csrc/aca.cuh — launch interface of the batched ACA on the device (aca.cu): SURVEY.md 8f rank 1, second step.
#ifndef HTB_ACA_CUH
#define HTB_ACA_CUH

#include "store.hpp"
#include <cuda_runtime.h>
#include <htool_b200.h>

namespace htb {

// Terms a block can hold, by team size: a block never needs more than floor(m n / (m + n)) terms (sympartialACA.hpp:102), i.e. 32 for the
// teams of 32 threads (max(m, n) <= 64) and 256 for the teams of 128 (max(m, n) <= 512): only blocks of the largest class can
// hit the limit (kAcaRankCap), when the reference itself would go past 512 terms.
HTB_HD constexpr int aca_max_rank(int team) { return team == 32 ? 32 : (team == 128 ? 256 : 512); }
HTB_HD constexpr int aca_team(int m, int n) { return (m > n ? m : n) > 512 ? 512 : ((m > n ? m : n) > 64 ? 128 : 32); }

// One admissible block of the block cluster tree (a leaf with rank == HTB_RANK_COMPRESS), in the root block's cluster
// numbering. "Dimension 1" is the rows when row_offset >= col_offset (GLOBAL offsets) and the columns otherwise
// (sympartialACA.hpp:46-66: the rule that makes the blocks (t, s) and (s, t) of a symmetric operator transposes of each other).
struct AcaBlock {
    int32_t lrow, lcol; // first row / column inside the root block
    int32_t m, n;
    uint32_t term_base; // first slot of the block in the term table
    uint16_t term_cap;  // slots of the block
    uint16_t swapped;   // 1: dimension 1 = columns
    uint32_t leaf;      // index of the leaf in the caller's descriptor
    uint32_t reserved;
};
static_assert(sizeof(AcaBlock) == 32, "AcaBlock must be 32 bytes");

// status of a block after the kernel (rank[] entry): q > 0 = rank; the others:
constexpr int kAcaFailed       = -1; // not advantageous / zero first row: the leaf becomes a dense leaf (tree_builder.hpp:617-626)
constexpr int kAcaPoolOverflow = -3; // the factor pool is full: the caller retries with a larger pool
constexpr int kAcaRankCap      = -4; // more than aca_max_rank(team) terms

// The factor pool: term j of a block is the chunk [uu_j (n1 coefficients) | vv_j (n2 coefficients)] at pool + 2 * term_off[term_base + j]
// (complex coefficients: re / im interleaved).
struct AcaPool {
    double *pool                = nullptr;
    unsigned long long capacity = 0; // doubles
    unsigned long long *cursor  = nullptr; // device counter, doubles
    uint32_t *term_off          = nullptr;
};

// Compresses blocks [first, first + count) with teams of `team` threads (32, 128 or 512) — one CTA per block.
// fma_axpy: the residual updates u -= c * v use one fused multiply-add (an FMA BLAS on the host) instead of a rounded
// product followed by a rounded sum (the SSE2 daxpy of the OpenBLAS this image carries).
// dots_mode: how the stopping criterion's dot products are computed by the teams of 128 / 512 threads — 0: by whole warps, with the
// in-order replay whenever the decision is close (aca.cu), 1: always in order (the reference's arithmetic, one thread per dot
// product), 2: warps + the replay at every iteration (tests). Teams of 32 threads always work in order.
cudaError_t launch_aca(int kernel, int team, const AcaBlock *blocks, long long first, long long count, const double *target_points, const double *source_points, double wavenumber, double epsilon, int fma_axpy, AcaPool pool,
                       int32_t *rank, int dots_mode, cudaStream_t st);

// Per leaf of the descriptor: where its factors are (LowRankTask.leaf indexes this table).
struct AcaLeaf {
    uint32_t term_base, n1, swapped, reserved;
};
// One warp per unit: copies the unit's panel out of the pool into the uploaded stream of side `side` (0: U panels, 1: V^T panels).
cudaError_t launch_scatter_lowrank(bool complex_coefficients, const DenseTask *tasks, long long n_tasks, int side, unsigned char *stream, const AcaLeaf *leaves, const double *pool, const uint32_t *term_off, cudaStream_t st);

} // namespace htb
#endif
