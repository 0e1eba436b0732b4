// htool_b200/csrc/generate.cu — leaf assembly on the device, first step: the dense near-field
// leaves are generated on the GPU straight into the side-0 stream of the leaf store.
//
// What it replaces in the reference: HMatrix::compute_dense_data (include/htool/hmatrix/hmatrix.hpp:222-226), one
// generator.copy_submatrix per dense leaf, called from HMatrixTreeBuilder::openmp_compute_blocks
// (hmatrix/tree_builder/tree_builder.hpp:629-633) or, in one batch, through the hook
// VirtualDenseBlocksGenerator::copy_dense_blocks (hmatrix/interfaces/virtual_dense_blocks_generator.hpp:12, called at
// tree_builder.hpp:650-665) — the shape of this entry point: a list of (rows, cols, row offset, col offset) blocks.
// At N = 1e6 that is 2.29 M leaves / 178.7 M coefficients which never exist on the host here.
//
// Built-in kernel functions = the analytic generators of the reference's test-suite
// (include/htool/testing/generator_test.hpp:155-205) plus the Helmholtz kernel of SURVEY.md 8d. Every operation is an
// explicitly rounded IEEE operation in the order the host generator performs it (no FMA contraction), so the real
// kernels are BIT-IDENTICAL to compute_dense_data; the Helmholtz kernel differs by the last ulps of sin / cos.
#include "generate.cuh"
#include "kernel_functions.cuh"

namespace htb {

namespace {

// One warp per dense unit.
template <bool CPLX, int KERNEL>
__global__ void generate_dense_kernel(const DenseTask *tasks, long long n_tasks, unsigned char *stream, const double *target_points, const double *source_points, double wavenumber) {
    const long long t = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tasks)
        return;
    const DenseTask task = tasks[t];
    const int lane       = threadIdx.x & 31;
    double *out          = reinterpret_cast<double *>(stream + task.byte_off);
    const bool diag      = (task.flags & (HTB_LEAF_DIAG_SYMMETRIC | HTB_LEAF_DIAG_HERMITIAN)) != 0;
    const bool upper     = (task.flags & HTB_LEAF_UPLO_UPPER) != 0;
    const bool herm      = (task.flags & HTB_LEAF_DIAG_HERMITIAN) != 0;
    const int total      = static_cast<int>(task.h) * task.w;
    for (int e = lane; e < total; e += 32) {
        const int k = e / task.h, i = e - k * task.h;
        int gi = task.p0 + i, gj = task.k0 + k;
        bool mirrored = false;
        if (diag && gi != gj && !(upper ? gi <= gj : gi >= gj)) { // symv / hemv read only the UPLO triangle: entry (j, i), conjugated for hemv
            const int tmp = gi;
            gi            = gj;
            gj            = tmp;
            mirrored      = true;
        }
        Value v = kernel_value<KERNEL>(target_points + 3ll * (task.lrow + gi), source_points + 3ll * (task.lcol + gj), wavenumber);
        if (herm && (mirrored || gi == gj))
            v.im = gi == gj ? 0. : -v.im; // hemv ignores the imaginary part of the diagonal
        const size_t at = static_cast<size_t>(k) * task.ld + i;
        if (CPLX) {
            out[2 * at]     = v.re;
            out[2 * at + 1] = v.im;
        } else
            out[at] = v.re;
    }
}

__global__ void scatter_headers_kernel(const StageDesc *stages, const unsigned long long *hdr_off, long long n_stages, const unsigned char *compact, unsigned char *stream) {
    const long long st = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (st >= n_stages)
        return;
    const int lane    = threadIdx.x & 31;
    const uint4 *src  = reinterpret_cast<const uint4 *>(compact + hdr_off[st]);
    uint4 *dst        = reinterpret_cast<uint4 *>(stream + stages[st].byte_off);
    const int n16     = static_cast<int>((hdr_off[st + 1] - hdr_off[st]) >> 4);
    for (int i = lane; i < n16; i += 32)
        dst[i] = src[i];
}

} // namespace

cudaError_t launch_scatter_headers(const StageDesc *stages, const unsigned long long *hdr_off, long long n_stages, const unsigned char *compact, unsigned char *stream, cudaStream_t st) {
    if (n_stages <= 0)
        return cudaSuccess;
    const int threads   = 256;
    const unsigned grid = static_cast<unsigned>((n_stages * 32 + threads - 1) / threads);
    scatter_headers_kernel<<<grid, threads, 0, st>>>(stages, hdr_off, n_stages, compact, stream);
    return cudaGetLastError();
}

bool kernel_is_complex(int kernel) { return kernel == HTB_KERNEL_COMPLEX_REG || kernel == HTB_KERNEL_HERMITIAN_REG || kernel == HTB_KERNEL_HELMHOLTZ || kernel == HTB_KERNEL_COMPLEX; }

cudaError_t launch_generate_dense(int kernel, const DenseTask *tasks, long long n_tasks, unsigned char *stream, const double *target_points, const double *source_points, double wavenumber, cudaStream_t st) {
    if (n_tasks == 0)
        return cudaSuccess;
    const int threads   = 256;
    const unsigned grid = static_cast<unsigned>((n_tasks * 32 + threads - 1) / threads);
    switch (kernel) {
    case HTB_KERNEL_LAPLACE:
        generate_dense_kernel<false, HTB_KERNEL_LAPLACE><<<grid, threads, 0, st>>>(tasks, n_tasks, stream, target_points, source_points, wavenumber);
        break;
    case HTB_KERNEL_LAPLACE_REG:
        generate_dense_kernel<false, HTB_KERNEL_LAPLACE_REG><<<grid, threads, 0, st>>>(tasks, n_tasks, stream, target_points, source_points, wavenumber);
        break;
    case HTB_KERNEL_COMPLEX_REG:
        generate_dense_kernel<true, HTB_KERNEL_COMPLEX_REG><<<grid, threads, 0, st>>>(tasks, n_tasks, stream, target_points, source_points, wavenumber);
        break;
    case HTB_KERNEL_HERMITIAN_REG:
        generate_dense_kernel<true, HTB_KERNEL_HERMITIAN_REG><<<grid, threads, 0, st>>>(tasks, n_tasks, stream, target_points, source_points, wavenumber);
        break;
    case HTB_KERNEL_HELMHOLTZ:
        generate_dense_kernel<true, HTB_KERNEL_HELMHOLTZ><<<grid, threads, 0, st>>>(tasks, n_tasks, stream, target_points, source_points, wavenumber);
        break;
    case HTB_KERNEL_COMPLEX:
        generate_dense_kernel<true, HTB_KERNEL_COMPLEX><<<grid, threads, 0, st>>>(tasks, n_tasks, stream, target_points, source_points, wavenumber);
        break;
    default:
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

} // namespace htb
