// htool_b200/csrc/mkernels.cuh — launch interface of the multi-RHS (FP64 tensor core) kernels (mkernels.cu).
#ifndef HTB_MKERNELS_CUH
#define HTB_MKERNELS_CUH

#include "kernels.cuh"

namespace htb {

// One pass over a side for a group of mc <= 64 REAL columns starting at real column col0 of ROW-major matrices. A
// complex<double> matrix is addressed through its real view (re / im interleaved): ld_in, ld_out, col0, mc, vs count
// doubles (twice the complex counts), rows stay rows.
struct MArgs {
    const double *in  = nullptr; // input matrix (REDUCE_M: multiplied rows; APPLY_M: rows of dense leaves, direction 0)
    long long in_rows = 0;
    int in_shift      = 0; // row index = block / leaf index + in_shift
    int ld_in         = 0; // mu
    double *out       = nullptr; // APPLY_M only
    long long out_rows = 0;
    int out_shift     = 0;
    int ld_out        = 0;
    int col0 = 0, mc = 0; // column group (col0: first real column of the group in `out`)
    int col0_in       = 0; // first real column of the group in `in` (0 when `in` is the padded staging copy of the group)
    int vs            = 0; // mc rounded up to a multiple of 8
    int vsp           = 0; // vector stride of the scratch, row stride of the B ring of APPLY_M and of the X block of REDUCE_M: vs + 4, i.e.
                           // = 4 (mod 8) doubles: the four rows k0 + tig of a DMMA fragment fall into four different groups of banks
                           // (a stride = 0 (mod 16) doubles puts them on the SAME banks: 4-way conflicts on every fragment load)
    double *mscratch  = nullptr; // one multi-RHS scratch copy: [TF | PARTM[0] | PARTM[1]] x vsp
    double alpha = 0., beta = 0.;
    double alpha_im = 0., beta_im = 0.; // complex<double> only
    int beta_is_zero = 0;
    int twice_only   = 0;
    int cplx         = 0; // coefficients are complex<double>
    int conj         = 0; // conjugate the coefficients (trans == 'C', Hermitian second application)
    int small_runs   = 1; // APPLY_M: run-private accumulators for runs of <= 5 row tiles per warp (0: the predicated per-tile walk)
    int reduce_split = 1; // REDUCE_M: jobs per 8-column tile of a tall run (groups of column tiles of the right-hand sides)
    int fast_tall    = 1; // APPLY_M: stage-level fast path for stages made of one full-height run of <= 32 columns
    int b_global     = 1; // APPLY_M: the B producers read the column tables from global memory (run ahead of the stage ring)
    int skip_dense   = 0; // APPLY_M: the dense columns of the runs (their tail, RunDesc::K_lr) are applied from the near-field panels instead
};

cudaError_t launch_reduce_m(const SideDevice &side, const LaunchConfig &cfg, const MArgs &args, cudaStream_t stream);
cudaError_t launch_apply_m(const SideDevice &side, const LaunchConfig &cfg, const MArgs &args, cudaStream_t stream);
// Sums the partials of the direction whose consumer is `side` into TF.
cudaError_t launch_combine_m(const SideDevice &side, double *mscratch, int vs, int vsp, int twice_only, cudaStream_t stream);
// Copies real columns [col0, col0 + mc) of the row-major matrix `in` (rows x ld_in) into dst (rows x vsp), zero padded: the
// staged group. Rows of the staged copy are one B-ring row apart, so the input rows of a dense leaf land with ONE bulk copy.
cudaError_t launch_stage_group(const double *in, long long rows, int ld_in, int col0, int mc, double *dst, int vsp, cudaStream_t stream);
// Fills the near-field panels (store.hpp, NearFieldLayout) from the dense units of the main stream of side 0.
cudaError_t launch_nf_copy(const NfTask *tasks, long long n_tasks, const unsigned char *src, unsigned char *dst, int esize, cudaStream_t stream);
size_t reduce_m_smem_bytes(const LaunchConfig &cfg, int vs, size_t esize);
size_t apply_m_smem_bytes(const LaunchConfig &cfg);
cudaError_t configure_mkernels(const LaunchConfig &cfg, size_t esize);

} // namespace htb
#endif
