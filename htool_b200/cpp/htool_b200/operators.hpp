// htool_b200/operators.hpp — the GPU twins of the reference's H-matrix operator adapters.
//
//   reference class (CPU)                                                       GPU twin (this file)
//   htool::RestrictedGlobalToLocalHMatrix  global_to_local_operators/hmatrix.hpp:15-36   htool_b200::RestrictedGlobalToLocalHMatrix
//   htool::LocalToLocalHMatrix             local_to_local_operators/hmatrix.hpp:15-56    htool_b200::LocalToLocalHMatrix
//
// Both implement the reference's plugin API unchanged — htool::VirtualGlobalToLocalOperator<T> /
// htool::VirtualLocalToLocalOperator<T> (distributed_operator/interfaces/*.hpp, three virtuals each) — so
// DistributedOperator::add_global_to_local_operator / add_local_to_local_operator
// (distributed_operator.hpp:47-53), CustomApproximationBuilder (distributed_operator/utility.hpp:22-35) and, through
// them, HPDDMOperator::GMV (wrappers/wrapper_hpddm.hpp:102-145) take them as they are.
//
// The global-to-local twin derives from the reference's own RestrictedGlobalToLocalOperator
// (restricted_operator.hpp:17-197): offsets, the optional in/out permutations, the beta pre-scaling of the
// global output for trans != 'N' and add_sub_matrix_product_to_local are therefore the reference's code, and
// only the two local_* hooks the reference forwards to openmp_internal_add_hmatrix_* are replaced.
#ifndef HTOOL_B200_OPERATORS_HPP
#define HTOOL_B200_OPERATORS_HPP

#include "device_hmatrix.hpp"
#include <algorithm>
#include <htool/distributed_operator/implementations/global_to_local_operators/restricted_operator.hpp>
#include <htool/distributed_operator/interfaces/virtual_local_to_local_operator.hpp>
#include <htool/distributed_operator/local_renumbering.hpp>
#include <vector>

namespace htool_b200 {

template <typename CoefficientPrecision, typename CoordinatePrecision = htool::underlying_type<CoefficientPrecision>>
class RestrictedGlobalToLocalHMatrix final : public htool::RestrictedGlobalToLocalOperator<CoefficientPrecision> {
    DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> m_data;

  public:
    /// Same arguments as the reference's constructor (global_to_local_operators/hmatrix.hpp:19) plus the CUDA
    /// device; the leaf store is uploaded here, once.
    RestrictedGlobalToLocalHMatrix(const htool::HMatrix<CoefficientPrecision, CoordinatePrecision> &hmatrix, const htool::LocalRenumbering &target_local_numbering, const htool::LocalRenumbering &source_local_numbering, bool target_use_permutation_to_mvprod = false, bool source_use_permutation_to_mvprod = false, int device = -1)
        : htool::RestrictedGlobalToLocalOperator<CoefficientPrecision>(target_local_numbering, source_local_numbering, target_use_permutation_to_mvprod, source_use_permutation_to_mvprod), m_data(hmatrix, device) {}

    void local_add_vector_product(char trans, CoefficientPrecision alpha, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out) const override {
        m_data.internal_add_vector_product(trans, alpha, in, beta, out);
    }
    void local_add_matrix_product_row_major(char trans, CoefficientPrecision alpha, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out, int mu) const override {
        m_data.internal_add_matrix_product_row_major(trans, alpha, in, beta, out, mu);
    }

    const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &get_device_hmatrix() const { return m_data; }
};

template <typename CoefficientPrecision, typename CoordinatePrecision = htool::underlying_type<CoefficientPrecision>>
class LocalToLocalHMatrix final : public htool::VirtualLocalToLocalOperator<CoefficientPrecision> {
    DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> m_data;

  public:
    explicit LocalToLocalHMatrix(const htool::HMatrix<CoefficientPrecision, CoordinatePrecision> &hmatrix, int device = -1) : m_data(hmatrix, device) {}

    void add_vector_product(char trans, CoefficientPrecision alpha, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out) const override {
        m_data.internal_add_vector_product(trans, alpha, in, beta, out);
    }
    void add_matrix_product_row_major(char trans, CoefficientPrecision alpha, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out, int mu) const override {
        m_data.internal_add_matrix_product_row_major(trans, alpha, in, beta, out, mu);
    }

    /// in = rows [offset, offset+size) of a global row-major matrix; out += H * (zero-extended in), alpha = beta = 1.
    /// Same clipping as the reference (local_to_local_operators/hmatrix.hpp:34-51).
    void add_sub_matrix_product_to_local(const CoefficientPrecision *const in, CoefficientPrecision *const out, int mu, int offset, int size) const override {
        const int source_offset = m_data.source_offset();
        const int source_size   = m_data.nb_cols();
        const int source_end    = source_size + source_offset;
        const int end           = size + offset;
        const int temp_offset   = std::max(offset, source_offset);
        const int temp_end      = std::min(source_end, end);
        if (offset == source_offset && temp_end == source_end) {
            add_matrix_product_row_major('N', 1, in, 1, out, mu);
        } else if (temp_end - temp_offset > 0) {
            std::vector<CoefficientPrecision> extension_by_zero(static_cast<size_t>(source_size) * mu, 0);
            std::copy_n(in + temp_offset - offset, static_cast<size_t>(temp_end - temp_offset) * mu, extension_by_zero.data() + static_cast<size_t>(temp_offset - source_offset) * mu);
            add_matrix_product_row_major('N', 1, extension_by_zero.data(), 1, out, mu);
        }
    }

    const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &get_device_hmatrix() const { return m_data; }
};

} // namespace htool_b200
#endif
