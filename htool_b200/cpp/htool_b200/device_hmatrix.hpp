// htool_b200/device_hmatrix.hpp — RAII owner of the device-resident leaf store of one htool::HMatrix.
//
// Host-side, header-only C++ (the reference is header-only C++ too). Everything below forwards to the C ABI
// of include/htool_b200.h; no CUDA header is needed to compile user code, only libhtool_b200.so to link.
//
// Error convention: the reference never throws and never returns codes from its products — it logs through
// htool::Logger and goes on (hmatrix/linalg/add_hmatrix_vector_product.hpp:112-115). The shim does the
// same: a failing C call is logged with the library's message (htb_last_error) and the call returns.
#ifndef HTOOL_B200_DEVICE_HMATRIX_HPP
#define HTOOL_B200_DEVICE_HMATRIX_HPP

#include "flatten.hpp"
#include <htool/hmatrix/interfaces/virtual_dense_blocks_generator.hpp>
#include <htool/hmatrix/interfaces/virtual_lrmat_generator.hpp>
#include <htool/misc/logger.hpp>
#include <string>
#include <unordered_set>
#include <vector>

namespace htool_b200 {

inline bool check(int status, const char *where) {
    if (status == HTB_OK) {
        return true;
    }
    // "Operation is not supported" is an ERROR in the reference (add_hmatrix_vector_product.hpp:113); anything
    // that stops the device from computing is CRITICAL, like the reference's "Missing permutation."
    htool::Logger::get_instance().log(status == HTB_ERR_UNSUPPORTED || status == HTB_ERR_INVALID ? htool::LogLevel::ERROR : htool::LogLevel::CRITICAL,
                                      std::string("[htool_b200] ") + where + ": " + htb_last_error()); // LCOV_EXCL_LINE
    return false;
}

/// RAII page-lock of a caller-owned host buffer (htb_host_register): while it lives, products given pointers inside
/// [ptr, ptr + n) touch the buffer directly from the kernels (single RHS: zero copy over PCIe) or by DMA without
/// staging. Meant for the vectors a Krylov solver reuses for every product of a solve (HPDDM's work vectors around
/// HPDDMOperator::GMV, wrappers/wrapper_hpddm.hpp:118-124).
template <typename T>
class PinnedHostBuffer {
    T *m_ptr = nullptr;

  public:
    PinnedHostBuffer(T *ptr, std::size_t n) {
        if (ptr != nullptr && n > 0 && check(htb_host_register(ptr, n * sizeof(T)), "htb_host_register")) {
            m_ptr = ptr;
        }
    }
    ~PinnedHostBuffer() {
        if (m_ptr != nullptr) {
            htb_host_unregister(m_ptr);
        }
    }
    PinnedHostBuffer(const PinnedHostBuffer &)            = delete;
    PinnedHostBuffer &operator=(const PinnedHostBuffer &) = delete;
    bool is_pinned() const { return m_ptr != nullptr; }
};

/// Leaf assembly on the device, first step. Plugged into the reference's builder with
/// HMatrixTreeBuilder::set_dense_blocks_generator (hmatrix/tree_builder/tree_builder.hpp:258), it receives the ONE batch
/// call the builder makes for all its dense leaves (copy_dense_blocks, tree_builder.hpp:650-665) and computes nothing: it
/// only records which host blocks were left unfilled. A DeviceHMatrix built with it (constructor below) generates those
/// leaves on the GPU, straight into the leaf store, from a built-in kernel function and the points.
/// The host H-matrix then holds zeros in those blocks: use it for assembly only, not for CPU products.
template <typename CoefficientPrecision>
class DeviceDenseBlocks final : public htool::VirtualDenseBlocksGenerator<CoefficientPrecision> {
    mutable std::unordered_set<const void *> m_blocks;

  public:
    void copy_dense_blocks(const std::vector<int> &, const std::vector<int> &, const std::vector<int> &, const std::vector<int> &, std::vector<CoefficientPrecision *> &ptr) const override {
        m_blocks.insert(ptr.begin(), ptr.end());
    }
    const std::unordered_set<const void *> &deferred() const { return m_blocks; }
    std::size_t size() const { return m_blocks.size(); }
};

/// Leaf assembly on the device, second step. Plugged into the reference's builder with
/// HMatrixTreeBuilder::set_low_rank_generator (hmatrix/tree_builder/tree_builder.hpp:251) in place of the default
/// sympartialACA (tree_builder.hpp:385), it receives the builder's call for every admissible block
/// (HMatrix::compute_low_rank_data, hmatrix.hpp:228-237) and computes nothing: it reports success and leaves the factors
/// empty. A DeviceHMatrix built with it (constructor below) compresses those blocks on the GPU with the reference's own
/// algorithm and stopping criterion (htb_create_compressed); blocks whose compression fails become dense leaves of the
/// device store, as in tree_builder.hpp:619-625. The host H-matrix then holds rank-0 leaves: use it for assembly only.
/// double and complex<double> (the reference's std::complex instantiation of sympartialACA).
/// A required rank (reqrank > 0) is not supported on the device: the call reports a failure and the builder computes the
/// block as a dense leaf on the host, which is what the reference does with any failed compression.
template <typename CoefficientPrecision>
class DeviceLowRankBlocks final : public htool::VirtualInternalLowRankGenerator<CoefficientPrecision> {
  public:
    bool copy_low_rank_approximation(int, int, int, int, htool::LowRankMatrix<CoefficientPrecision> &) const override { return true; }
    bool copy_low_rank_approximation(int, int, int, int, int, htool::LowRankMatrix<CoefficientPrecision> &) const override {
        htool::Logger::get_instance().log(htool::LogLevel::WARNING, "[htool_b200] a required rank is not supported by the device compression: dense block instead"); // LCOV_EXCL_LINE
        return false;
    }
};

/// A built-in kernel function (htb_generator_desc) with the geometry in USER numbering, as a VirtualGenerator sees it.
struct BuiltinKernel {
    int kernel        = HTB_KERNEL_LAPLACE_REG;
    double wavenumber = 0.;
    const double *target_points = nullptr; // 3 x number of target points
    const double *source_points = nullptr;
};

/// The leaf store of `hmatrix` on one B200: flattened with flatten() and uploaded once, at construction
/// (north_star item 1). The HMatrix can be destroyed afterwards: nothing on the host is referenced again.
template <typename CoefficientPrecision, typename CoordinatePrecision = htool::underlying_type<CoefficientPrecision>>
class DeviceHMatrix {
    htb_handle m_handle = nullptr;
    int m_nb_rows = 0, m_nb_cols = 0, m_target_offset = 0, m_source_offset = 0;
    int64_t m_nb_leaves = 0;

  public:
    explicit DeviceHMatrix(const htool::HMatrix<CoefficientPrecision, CoordinatePrecision> &hmatrix, int device = -1) {
        FlatHMatrix flat = flatten(hmatrix, device);
        m_nb_rows        = flat.desc.nb_rows;
        m_nb_cols        = flat.desc.nb_cols;
        m_target_offset  = flat.desc.row_offset;
        m_source_offset  = flat.desc.col_offset;
        if (!check(htb_create(&flat.desc, &m_handle), "htb_create")) {
            m_handle = nullptr;
            return;
        }
        // permutations restricted to the root block, shifted to [0, size): what user_to_cluster / cluster_to_user
        // index with (clustering/cluster_node.hpp:150-175)
        const auto &tp = hmatrix.get_target_cluster().get_permutation();
        const auto &sp = hmatrix.get_source_cluster().get_permutation();
        std::vector<int32_t> t(m_nb_rows), s(m_nb_cols);
        bool local = true;
        for (int i = 0; i < m_nb_rows; i++) {
            t[i] = tp[m_target_offset + i] - m_target_offset;
            local = local && t[i] >= 0 && t[i] < m_nb_rows;
        }
        for (int i = 0; i < m_nb_cols; i++) {
            s[i] = sp[m_source_offset + i] - m_source_offset;
            local = local && s[i] >= 0 && s[i] < m_nb_cols;
        }
        if (local) { // a sub-block of a non-local partition has no block-local permutation: user-numbering products are then unavailable, as in the reference
            check(htb_set_permutations(m_handle, t.data(), s.data()), "htb_set_permutations");
        }
    }
    /// Device-side assembly of the dense leaves: `hmatrix` was built with `deferred` as its dense-blocks generator, so its
    /// dense leaves hold no coefficients; they are generated on the GPU from `kernel` (htb_create_generated). Low-rank
    /// leaves (and admissible blocks whose compression failed) are packed from the host as usual.
    DeviceHMatrix(const htool::HMatrix<CoefficientPrecision, CoordinatePrecision> &hmatrix, const DeviceDenseBlocks<CoefficientPrecision> &deferred, const BuiltinKernel &kernel, int device = -1) {
        assemble_on_device(hmatrix, deferred, kernel, false, device);
    }
    /// Device-side assembly of ALL the leaves: `hmatrix` was built with `deferred` as its dense-blocks generator and a
    /// DeviceLowRankBlocks as its low-rank generator, so it is a block cluster tree without coefficients. The admissible
    /// blocks are compressed on the GPU by the reference's sympartialACA at the builder's epsilon, the dense leaves are
    /// generated (htb_create_compressed). Any built-in kernel function, double or complex<double>.
    DeviceHMatrix(const htool::HMatrix<CoefficientPrecision, CoordinatePrecision> &hmatrix, const DeviceDenseBlocks<CoefficientPrecision> &deferred, const DeviceLowRankBlocks<CoefficientPrecision> &, const BuiltinKernel &kernel, int device = -1) {
        assemble_on_device(hmatrix, deferred, kernel, true, device);
    }
    /// Ranks found by the device compression, in the order of htool::get_leaves_from (-1: dense leaf).
    std::vector<int32_t> leaf_ranks() const {
        std::vector<int32_t> r(static_cast<std::size_t>(m_nb_leaves));
        if (!m_handle || !check(htb_get_leaf_ranks(m_handle, r.data(), m_nb_leaves), "htb_get_leaf_ranks")) {
            r.clear();
        }
        return r;
    }

  private:
    void assemble_on_device(const htool::HMatrix<CoefficientPrecision, CoordinatePrecision> &hmatrix, const DeviceDenseBlocks<CoefficientPrecision> &deferred, const BuiltinKernel &kernel, bool compress, int device) {
        double epsilon   = -1.;
        FlatHMatrix flat = flatten(hmatrix, device, &deferred.deferred(), compress, &epsilon);
        m_nb_leaves      = flat.desc.nb_leaves;
        m_nb_rows        = flat.desc.nb_rows;
        m_nb_cols        = flat.desc.nb_cols;
        m_target_offset  = flat.desc.row_offset;
        m_source_offset  = flat.desc.col_offset;
        // the points of the root block in cluster numbering: row i <-> user index permutation[offset + i]
        const auto &tp = hmatrix.get_target_cluster().get_permutation();
        const auto &sp = hmatrix.get_source_cluster().get_permutation();
        std::vector<double> tpts(static_cast<std::size_t>(3) * m_nb_rows), spts(static_cast<std::size_t>(3) * m_nb_cols);
        for (int i = 0; i < m_nb_rows; i++) {
            for (int d = 0; d < 3; d++) {
                tpts[3 * static_cast<std::size_t>(i) + d] = kernel.target_points[3 * static_cast<std::size_t>(tp[m_target_offset + i]) + d];
            }
        }
        for (int i = 0; i < m_nb_cols; i++) {
            for (int d = 0; d < 3; d++) {
                spts[3 * static_cast<std::size_t>(i) + d] = kernel.source_points[3 * static_cast<std::size_t>(sp[m_source_offset + i]) + d];
            }
        }
        htb_generator_desc gen{};
        gen.kernel            = kernel.kernel;
        gen.spatial_dimension = 3;
        gen.wavenumber        = kernel.wavenumber;
        gen.target_points     = tpts.data();
        gen.source_points     = spts.data();
        const bool ok = compress && epsilon > 0. ? check(htb_create_compressed(&flat.desc, &gen, epsilon, &m_handle), "htb_create_compressed") : check(htb_create_generated(&flat.desc, &gen, &m_handle), "htb_create_generated");
        if (!ok) {
            m_handle = nullptr;
        }
    }

  public:
    DeviceHMatrix(const DeviceHMatrix &)            = delete;
    DeviceHMatrix &operator=(const DeviceHMatrix &) = delete;
    DeviceHMatrix(DeviceHMatrix &&o) noexcept : m_handle(o.m_handle), m_nb_rows(o.m_nb_rows), m_nb_cols(o.m_nb_cols), m_target_offset(o.m_target_offset), m_source_offset(o.m_source_offset), m_nb_leaves(o.m_nb_leaves) { o.m_handle = nullptr; }
    DeviceHMatrix &operator=(DeviceHMatrix &&o) noexcept {
        if (this != &o) {
            if (m_handle) {
                htb_destroy(m_handle);
            }
            m_handle        = o.m_handle;
            m_nb_rows       = o.m_nb_rows;
            m_nb_cols       = o.m_nb_cols;
            m_target_offset = o.m_target_offset;
            m_source_offset = o.m_source_offset;
            m_nb_leaves     = o.m_nb_leaves;
            o.m_handle      = nullptr;
        }
        return *this;
    }
    ~DeviceHMatrix() {
        if (m_handle) {
            htb_destroy(m_handle);
        }
    }

    bool is_valid() const { return m_handle != nullptr; }
    htb_handle get() const { return m_handle; }
    int nb_rows() const { return m_nb_rows; }
    int nb_cols() const { return m_nb_cols; }
    int target_offset() const { return m_target_offset; }
    int source_offset() const { return m_source_offset; }
    htb_info info() const {
        htb_info i{};
        if (m_handle) {
            check(htb_get_info(m_handle, &i), "htb_get_info");
        }
        return i;
    }

    // ---- cluster numbering, host pointers: the two calls the operator adapters make ---------------------------
    /// Same contract as openmp_internal_add_hmatrix_vector_product (add_hmatrix_vector_product.hpp:107-170).
    void internal_add_vector_product(char trans, CoefficientPrecision alpha, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out) const {
        if (!m_handle) {
            htool::Logger::get_instance().log(htool::LogLevel::CRITICAL, "[htool_b200] no device leaf store: product skipped"); // LCOV_EXCL_LINE
            return;
        }
        check(htb_add_vector_product(m_handle, trans, &alpha, in, &beta, out, HTB_MEM_HOST), "htb_add_vector_product");
    }
    /// Same contract as openmp_internal_add_hmatrix_matrix_product_row_major(trans,'N',...) (add_hmatrix_matrix_product_row_major.hpp:112-178).
    void internal_add_matrix_product_row_major(char trans, CoefficientPrecision alpha, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out, int mu) const {
        if (!m_handle) {
            htool::Logger::get_instance().log(htool::LogLevel::CRITICAL, "[htool_b200] no device leaf store: product skipped"); // LCOV_EXCL_LINE
            return;
        }
        check(htb_add_matrix_product_row_major(m_handle, trans, &alpha, in, &beta, out, mu, HTB_MEM_HOST), "htb_add_matrix_product_row_major");
    }
};

// ---- free functions with the reference's names and argument meaning ---------------------------------------------
// (hmatrix/linalg/add_hmatrix_vector_product.hpp:173-197, add_hmatrix_matrix_product.hpp:176-205). The execution-policy
// argument of the reference is accepted and ignored: there is one way to run on the device.

/// y <- alpha op(H) x + beta y in USER numbering; the permutation gather/scatter run on the device.
template <typename CoefficientPrecision, typename CoordinatePrecision>
void add_hmatrix_vector_product(char trans, CoefficientPrecision alpha, const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &A, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out) {
    if (A.is_valid()) {
        check(htb_add_vector_product_user_numbering(A.get(), trans, &alpha, in, &beta, out, HTB_MEM_HOST), "htb_add_vector_product_user_numbering");
    }
}
template <typename ExecutionPolicy, typename CoefficientPrecision, typename CoordinatePrecision>
void add_hmatrix_vector_product(ExecutionPolicy &&, char trans, CoefficientPrecision alpha, const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &A, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out, CoefficientPrecision * = nullptr) {
    add_hmatrix_vector_product(trans, alpha, A, in, beta, out);
}

/// C <- alpha op(H) B + beta C, B and C COLUMN-major with mu columns, user numbering (transb must be 'N',
/// as in the row-major kernels of the reference, add_hmatrix_matrix_product_row_major.hpp:121-123).
template <typename CoefficientPrecision, typename CoordinatePrecision>
void add_hmatrix_matrix_product(char transa, char transb, CoefficientPrecision alpha, const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &A, const CoefficientPrecision *B, CoefficientPrecision beta, CoefficientPrecision *C, int mu) {
    if (transb != 'N') {
        htool::Logger::get_instance().log(htool::LogLevel::ERROR, "Operation is not supported (transb=" + std::string(1, transb) + ")"); // LCOV_EXCL_LINE
        return;
    }
    if (A.is_valid()) {
        check(htb_add_matrix_product_user_numbering(A.get(), transa, &alpha, B, &beta, C, mu, HTB_MEM_HOST), "htb_add_matrix_product_user_numbering");
    }
}
template <typename ExecutionPolicy, typename CoefficientPrecision, typename CoordinatePrecision>
void add_hmatrix_matrix_product(ExecutionPolicy &&, char transa, char transb, CoefficientPrecision alpha, const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &A, const CoefficientPrecision *B, CoefficientPrecision beta, CoefficientPrecision *C, int mu, CoefficientPrecision * = nullptr) {
    add_hmatrix_matrix_product(transa, transb, alpha, A, B, beta, C, mu);
}

/// Cluster numbering, same names as the reference's internal kernels.
template <typename CoefficientPrecision, typename CoordinatePrecision>
void internal_add_hmatrix_vector_product(char trans, CoefficientPrecision alpha, const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &A, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out) {
    A.internal_add_vector_product(trans, alpha, in, beta, out);
}
template <typename CoefficientPrecision, typename CoordinatePrecision>
void internal_add_hmatrix_matrix_product_row_major(char transa, char transb, CoefficientPrecision alpha, const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &A, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out, int mu) {
    if (transb != 'N') {
        htool::Logger::get_instance().log(htool::LogLevel::ERROR, "Operation is not implemented for sequential_internal_add_hmatrix_matrix_product_row_major (transb=" + std::string(1, transb) + ")"); // LCOV_EXCL_LINE
        return;
    }
    A.internal_add_matrix_product_row_major(transa, alpha, in, beta, out, mu);
}

} // namespace htool_b200
#endif
