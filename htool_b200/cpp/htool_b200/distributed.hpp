// htool_b200/distributed.hpp — the row-sharded product over the GPUs of one box, behind the reference's names.
//
// The reference distributes an operator as one row strip per MPI rank (DistributedOperator,
// distributed_operator/distributed_operator.hpp:18-71) and moves vectors with MPI collectives on HOST memory
// (linalg/utility.hpp:11-28, add_distributed_operator_vector_product_local_to_local.hpp:19-89,
// add_distributed_operator_vector_product_global_to_global.hpp:18-85). With the GPU twins of operators.hpp
// plugged into a DistributedOperator those functions keep working unchanged — every product then crosses
// PCIe twice and the exchange stays on the host.
//
// DeviceDistributedOperator is the alternative for one process per GPU: same strip (built exactly as
// DefaultApproximationBuilder does, distributed_operator/utility.hpp:56), same partition object, same MPI
// communicator for the bootstrap only (one MPI_Bcast of the 128-byte NCCL id); afterwards the exchange runs over
// NVLink: for 'N' a push kernel stores the rank's slice of x into every peer's buffer (CUDA IPC peer mappings) and the
// product's first pass waits block by block for the slices it needs; T / C and global-to-global use NCCL
// (htb_dist_add_product_local_to_local / htb_dist_add_product_global_to_global, htool_b200/csrc/dist.cu). The free functions below have
// the reference's names and argument meaning so that call sites (HPDDMOperator::GMV,
// wrappers/wrapper_hpddm.hpp:118-124) only change the operator type; `work` is accepted and ignored.
#ifndef HTOOL_B200_DISTRIBUTED_HPP
#define HTOOL_B200_DISTRIBUTED_HPP

#include "device_hmatrix.hpp"
#include <htool/distributed_operator/interfaces/virtual_partition.hpp>
#include <mpi.h>
#include <utility>
#include <vector>

namespace htool_b200 {

template <typename CoefficientPrecision, typename CoordinatePrecision = htool::underlying_type<CoefficientPrecision>>
class DeviceDistributedOperator {
    DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> m_data;
    const htool::VirtualPartition<CoefficientPrecision> &m_partition;
    MPI_Comm m_comm;
    int m_rank = 0, m_size = 1;
    bool m_ready = false;

    /// true iff `local_ok` holds on EVERY rank of the communicator
    bool agree(bool local_ok) const {
        int mine = local_ok ? 1 : 0, all = 0;
        MPI_Allreduce(&mine, &all, 1, MPI_INT, MPI_MIN, m_comm);
        return all == 1;
    }

    /// The bootstrap is failure-COLLECTIVE: every rank takes part in every collective below whatever happened to it
    /// locally, and the ranks agree (MPI_Allreduce MIN) on success after each step, so that no rank is ever left alone
    /// inside MPI_Bcast / ncclCommInitRank / a later NCCL call or flag wait (the reference never leaves a rank inside a
    /// collective alone either: its products have no per-rank early exit).
    bool bootstrap() {
        unsigned char id[HTB_NCCL_UNIQUE_ID_BYTES] = {0};
        bool ok = m_data.is_valid();
        if (m_rank == 0) {
            ok = check(htb_nccl_get_unique_id(id), "htb_nccl_get_unique_id") && ok;
        }
        MPI_Bcast(id, HTB_NCCL_UNIQUE_ID_BYTES, MPI_UNSIGNED_CHAR, 0, m_comm);
        if (!agree(ok)) {
            return false;
        }
        std::vector<int32_t> offsets(m_size + 1, 0);
        for (int r = 0; r < m_size; r++) {
            offsets[r]     = m_partition.get_offset_of_partition(r);
            offsets[r + 1] = offsets[r] + m_partition.get_size_of_partition(r);
        }
        ok             = check(htb_comm_init(m_data.get(), id, m_size, m_rank, offsets.data()), "htb_comm_init");
        const bool all = agree(ok);
        if (!all && ok) {
            htb_comm_destroy(m_data.get());
        }
        return all;
    }

  public:
    /// `strip` = this rank's block row, HMatrixTreeBuilder::build(generator, target, source, rank, rank)
    /// (distributed_operator/utility.hpp:56); `partition` = the partition of BOTH the target and the source
    /// cluster tree (square operator, as in DefaultApproximationBuilder's symmetric constructor, utility.hpp:61).
    DeviceDistributedOperator(const htool::HMatrix<CoefficientPrecision, CoordinatePrecision> &strip, const htool::VirtualPartition<CoefficientPrecision> &partition, MPI_Comm comm, int device = -1)
        : m_data(strip, device), m_partition(partition), m_comm(comm) {
        MPI_Comm_rank(comm, &m_rank);
        MPI_Comm_size(comm, &m_size);
        m_ready = bootstrap();
    }
    /// Same, for a strip whose leaf store already exists — e.g. one ASSEMBLED ON THE DEVICE from the block cluster tree
    /// (DeviceHMatrix(tree, dense, lowrank, kernel), device_hmatrix.hpp): every rank compresses its own block row on its own GPU.
    DeviceDistributedOperator(DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &&strip, const htool::VirtualPartition<CoefficientPrecision> &partition, MPI_Comm comm)
        : m_data(std::move(strip)), m_partition(partition), m_comm(comm) {
        MPI_Comm_rank(comm, &m_rank);
        MPI_Comm_size(comm, &m_size);
        m_ready = bootstrap();
    }
    ~DeviceDistributedOperator() {
        if (m_ready) {
            htb_comm_destroy(m_data.get());
        }
    }
    DeviceDistributedOperator(const DeviceDistributedOperator &)            = delete;
    DeviceDistributedOperator &operator=(const DeviceDistributedOperator &) = delete;

    bool is_valid() const { return m_ready; }
    MPI_Comm get_comm() const { return m_comm; }
    const htool::VirtualPartition<CoefficientPrecision> &get_target_partition() const { return m_partition; }
    const htool::VirtualPartition<CoefficientPrecision> &get_source_partition() const { return m_partition; }
    const DeviceHMatrix<CoefficientPrecision, CoordinatePrecision> &get_device_hmatrix() const { return m_data; }

    void local_to_local(char trans, CoefficientPrecision alpha, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out, int mu) const {
        if (m_ready) {
            check(htb_dist_add_product_local_to_local(m_data.get(), trans, &alpha, in, &beta, out, mu, HTB_MEM_HOST), "htb_dist_add_product_local_to_local");
        }
    }
    /// Unpreconditioned restarted GMRES with the Krylov loop resident on the device (htb_gmres): the counterpart of
    /// DDM::solve with "-hpddm_schwarz_method none" (solvers/ddm.hpp:134-193). rhs / x: this rank's slices in partition
    /// numbering; x = initial guess on entry. options == nullptr: HPDDM's defaults (restart 40, 100 iterations, 1e-6).
    htb_gmres_result solve(const CoefficientPrecision *rhs, CoefficientPrecision *x, const htb_gmres_options *options = nullptr) const {
        htb_gmres_result result{};
        result.true_relative_residual = -1.;
        if (m_ready) {
            check(htb_gmres(m_data.get(), rhs, x, options, &result, HTB_MEM_HOST), "htb_gmres");
        }
        return result;
    }
    void global_to_global(char trans, CoefficientPrecision alpha, const CoefficientPrecision *in, CoefficientPrecision beta, CoefficientPrecision *out, int mu) const {
        if (m_ready) {
            check(htb_dist_add_product_global_to_global(m_data.get(), trans, &alpha, in, &beta, out, mu, HTB_MEM_HOST), "htb_dist_add_product_global_to_global");
        }
    }
};

// ---- the reference's entry points, overloaded on the device operator (partition numbering, host pointers) --------

/// add_distributed_operator_vector_product_local_to_local.hpp:19 — what HPDDMOperator::GMV calls per Krylov iteration.
template <typename T, typename U>
void internal_add_distributed_operator_vector_product_local_to_local(char trans, T alpha, const DeviceDistributedOperator<T, U> &A, const T *const in, T beta, T *const out, T * /*work*/ = nullptr) {
    A.local_to_local(trans, alpha, in, beta, out, 1);
}
/// add_distributed_operator_matrix_product_row_major_local_to_local.hpp:25 with raw row-major pointers (mu contiguous).
template <typename T, typename U>
void internal_add_distributed_operator_matrix_product_row_major_local_to_local(char trans, T alpha, const DeviceDistributedOperator<T, U> &A, const T *const in, T beta, T *const out, int mu, T * /*work*/ = nullptr) {
    A.local_to_local(trans, alpha, in, beta, out, mu);
}
/// add_distributed_operator_vector_product_global_to_global.hpp:18 (partition numbering in and out).
template <typename T, typename U>
void internal_add_distributed_operator_vector_product_global_to_global(char trans, T alpha, const DeviceDistributedOperator<T, U> &A, const T *const in, T beta, T *const out, T * /*work*/ = nullptr) {
    A.global_to_global(trans, alpha, in, beta, out, 1);
}
/// add_distributed_operator_matrix_product_row_major_global_to_global.hpp:18 with raw row-major pointers.
template <typename T, typename U>
void internal_add_distributed_operator_matrix_product_row_major_global_to_global(char trans, T alpha, const DeviceDistributedOperator<T, U> &A, const T *const in, T beta, T *const out, int mu, T * /*work*/ = nullptr) {
    A.global_to_global(trans, alpha, in, beta, out, mu);
}
/// add_distributed_operator_vector_product_global_to_global.hpp:97-118: user numbering around the internal product,
/// with the partition's own renumbering (PartitionFromCluster::global_to_partition_numbering, partition_from_cluster.hpp:27-32).
template <typename T, typename U>
void add_distributed_operator_vector_product_global_to_global(char trans, T alpha, const DeviceDistributedOperator<T, U> &A, const T *const in, T beta, T *const out, T * /*work*/ = nullptr) {
    const auto &partition = A.get_source_partition();
    const int n           = partition.get_global_size();
    std::vector<T> in_p(n), out_p(n);
    partition.global_to_partition_numbering(in, in_p.data());
    if (beta != T(0)) {
        partition.global_to_partition_numbering(out, out_p.data());
    }
    A.global_to_global(trans, alpha, in_p.data(), beta, out_p.data(), 1);
    partition.partition_to_global_numbering(out_p.data(), out);
}

} // namespace htool_b200
#endif
