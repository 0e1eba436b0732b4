// htool_b200/flatten.hpp — walk an htool::HMatrix once and describe its leaves for the C ABI.
//
// Host-side, header-only, compiled by the user next to htool's own headers. It reads the reference's
// objects and never re-derives anything: offsets/sizes come from Cluster::get_offset/get_size
// (clustering/cluster_node.hpp:55-56), the leaf lists from get_leaves_from (hmatrix/hmatrix.hpp:248-274),
// payload pointers from Matrix::data() (matrix/matrix.hpp) and LowRankMatrix::get_U/get_V
// (hmatrix/lrmat/lrmat.hpp:43-54). Block-tree indexing is therefore bit-exact by construction.
#ifndef HTOOL_B200_FLATTEN_HPP
#define HTOOL_B200_FLATTEN_HPP

#include <htool_b200.h>
#include <complex>
#include <htool/hmatrix/hmatrix.hpp>
#include <type_traits>
#include <unordered_set>
#include <vector>

namespace htool_b200 {

template <typename T>
struct coefficient_type;
template <>
struct coefficient_type<double> {
    static constexpr int value = HTB_DOUBLE;
};
template <>
struct coefficient_type<std::complex<double>> {
    static constexpr int value = HTB_COMPLEX_DOUBLE;
};

/// Leaf descriptors + the root description, in the order get_leaves_from returns the leaves.
struct FlatHMatrix {
    std::vector<htb_leaf> leaves;
    htb_hmatrix_desc desc{};
};

/// `deferred_dense` (optional): dense blocks whose coefficients were NOT computed on the host (DeviceDenseBlocks below,
/// device_hmatrix.hpp); their leaves get data0 = nullptr and are generated on the device by htb_create_generated.
/// `compress_on_device`: the low-rank leaves were "computed" by DeviceLowRankBlocks (device_hmatrix.hpp), i.e. not at all: they
/// get rank = HTB_RANK_COMPRESS and no data, and are compressed on the device by htb_create_compressed; `epsilon_out` receives
/// the tolerance the builder gave them (LowRankMatrix::get_epsilon, what sympartialACA reads, sympartialACA.hpp:32).
template <typename CoefficientPrecision, typename CoordinatePrecision = htool::underlying_type<CoefficientPrecision>>
FlatHMatrix flatten(const htool::HMatrix<CoefficientPrecision, CoordinatePrecision> &hmatrix, int device = -1, const std::unordered_set<const void *> *deferred_dense = nullptr, bool compress_on_device = false, double *epsilon_out = nullptr) {
    using HMatrixType = htool::HMatrix<CoefficientPrecision, CoordinatePrecision>;
    FlatHMatrix flat;

    // Same call the reference makes at the top of every product (add_hmatrix_vector_product.hpp:110).
    std::vector<const HMatrixType *> leaves, leaves_for_symmetry;
    std::tie(leaves, leaves_for_symmetry) = htool::get_leaves_from(hmatrix);
    std::unordered_set<const HMatrixType *> applied_twice(leaves_for_symmetry.begin(), leaves_for_symmetry.end());

    const int root_target_offset = hmatrix.get_target_cluster().get_offset();
    const int root_source_offset = hmatrix.get_source_cluster().get_offset();

    flat.leaves.reserve(leaves.size());
    for (const HMatrixType *leaf : leaves) {
        htb_leaf d{};
        d.row_offset = leaf->get_target_cluster().get_offset() - root_target_offset;
        d.col_offset = leaf->get_source_cluster().get_offset() - root_source_offset;
        d.nb_rows    = leaf->get_target_cluster().get_size();
        d.nb_cols    = leaf->get_source_cluster().get_size();
        d.flags      = 0;
        if (applied_twice.count(leaf)) {
            d.flags |= HTB_LEAF_APPLY_TRANSPOSED_TOO;
        }
        if (leaf->is_dense()) {
            d.rank  = -1;
            d.data0 = leaf->get_dense_data()->data();
            d.data1 = nullptr;
            if (deferred_dense != nullptr && deferred_dense->count(d.data0) != 0) {
                d.data0 = nullptr; // the host block was allocated but never filled
            }
            // Diagonal dense leaves go through symv/hemv (add_hmatrix_vector_product.hpp:22-24,41-45).
            if (leaf->get_symmetry() == 'S') {
                d.flags |= HTB_LEAF_DIAG_SYMMETRIC;
            } else if (leaf->get_symmetry() == 'H') {
                d.flags |= HTB_LEAF_DIAG_HERMITIAN;
            }
            if (leaf->get_symmetry() != 'N' && leaf->get_UPLO() == 'U') {
                d.flags |= HTB_LEAF_UPLO_UPPER;
            }
        } else if (leaf->is_low_rank()) {
            const auto *lrmat = leaf->get_low_rank_data();
            d.rank            = lrmat->rank_of();
            d.data0           = lrmat->get_U().data();
            d.data1           = lrmat->get_V().data();
            if (compress_on_device && d.rank == 0) {
                d.rank  = HTB_RANK_COMPRESS;
                d.data0 = nullptr;
                d.data1 = nullptr;
                if (epsilon_out != nullptr) {
                    *epsilon_out = static_cast<double>(lrmat->get_epsilon());
                }
            }
        } else {
            continue; // a childless hierarchical node holds no data and contributes nothing
        }
        flat.leaves.push_back(d);
    }

    flat.desc.dtype               = coefficient_type<CoefficientPrecision>::value;
    flat.desc.nb_rows             = hmatrix.get_target_cluster().get_size();
    flat.desc.nb_cols             = hmatrix.get_source_cluster().get_size();
    flat.desc.row_offset          = root_target_offset;
    flat.desc.col_offset          = root_source_offset;
    flat.desc.symmetry_for_leaves = hmatrix.get_symmetry_for_leaves();
    flat.desc.uplo_for_leaves     = hmatrix.get_UPLO_for_leaves();
    flat.desc.device              = device;
    flat.desc.nb_leaves           = static_cast<int64_t>(flat.leaves.size());
    flat.desc.leaves              = flat.leaves.data();
    return flat;
}

} // namespace htool_b200
#endif
