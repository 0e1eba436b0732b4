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
    int col0 = 0, mc = 0; // column group
    int vs            = 0; // mc rounded up to a multiple of 8
    int vsp           = 0; // vector stride of the scratch (= vs) and row stride of the B ring of APPLY_M: consecutive vectors of a piece are consecutive rows
    double *mscratch  = nullptr; // one multi-RHS scratch copy: [TF | PARTM[0] | PARTM[1]] x vsp
    double alpha = 0., beta = 0.;
    double alpha_im = 0., beta_im = 0.; // complex<double> only
    int beta_is_zero = 0;
    int twice_only   = 0;
    int cplx         = 0; // coefficients are complex<double>
    int conj         = 0; // conjugate the coefficients (trans == 'C', Hermitian second application)
};

cudaError_t launch_reduce_m(const SideDevice &side, const LaunchConfig &cfg, const MArgs &args, cudaStream_t stream);
cudaError_t launch_apply_m(const SideDevice &side, const LaunchConfig &cfg, const MArgs &args, cudaStream_t stream);
// Sums the partials of the direction whose consumer is `side` into TF.
cudaError_t launch_combine_m(const SideDevice &side, double *mscratch, int vs, int vsp, int twice_only, cudaStream_t stream);
size_t reduce_m_smem_bytes(const LaunchConfig &cfg, int vs, size_t esize);
size_t apply_m_smem_bytes(const LaunchConfig &cfg);
cudaError_t configure_mkernels(const LaunchConfig &cfg, size_t esize);

} // namespace htb
#endif
