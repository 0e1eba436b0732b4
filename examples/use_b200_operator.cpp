// examples/use_b200_operator.cpp — the flow of the reference's examples/use_hmatrix.cpp and
// examples/use_distributed_operator.cpp with the B200 product plugged in. Htool (unmodified, header-only) still does the
// clustering, the block tree and the ACA compression on the host; the three ways of using the device product are shown:
//
//   (1) user numbering, single process:   htool_b200::DeviceHMatrix + add_hmatrix_vector_product   (use_hmatrix.cpp:107)
//   (2) Htool's own DistributedOperator with the GPU twin of RestrictedGlobalToLocalHMatrix registered through
//       CustomApproximationBuilder (distributed_operator/utility.hpp:22-35): every reference entry point keeps working,
//       HPDDM included (use_distributed_operator.cpp:108);
//   (3) one process per GPU with the exchange on NVLink: htool_b200::DeviceDistributedOperator, and the device-resident
//       GMRES that stands for DDM::solve with -hpddm_schwarz_method none;
//   (4) the leaf assembly itself on the GPU (batched sympartialACA + dense generation): the host keeps the block tree only.
//
// Build:
//   g++ -std=c++17 -fopenmp examples/use_b200_operator.cpp -I<htool>/include -Iinclude -Ihtool_b200/cpp \
//       -Lhtool_b200/lib -lhtool_b200 -Wl,-rpath,$PWD/htool_b200/lib <BLAS / LAPACK / MPI as Htool needs>
#include <htool/clustering/tree_builder/tree_builder.hpp>
#include <htool/distributed_operator/distributed_operator.hpp>
#include <htool/distributed_operator/implementations/partition_from_cluster.hpp>
#include <htool/distributed_operator/linalg.hpp>
#include <htool/distributed_operator/utility.hpp>
#include <htool/hmatrix/tree_builder/tree_builder.hpp>
#include <htool/testing/geometry.hpp>

#include <htool_b200/distributed.hpp>
#include <htool_b200/operators.hpp>

#include <cmath>
#include <memory>
#include <iostream>
#include <vector>

// 1 / (1e-5 + 4 pi r): the regularised Laplace kernel of the benchmark, symmetric, finite on the diagonal
class LaplaceKernel final : public htool::VirtualGenerator<double> {
    const std::vector<double> &m_points;

  public:
    explicit LaplaceKernel(const std::vector<double> &points) : m_points(points) {}
    double entry(int i, int j) const {
        double r2 = 0;
        for (int d = 0; d < 3; d++) {
            const double diff = m_points[3 * i + d] - m_points[3 * j + d];
            r2 += diff * diff;
        }
        return 1. / (1e-5 + 4. * M_PI * std::sqrt(r2));
    }
    void copy_submatrix(int M, int N, const int *rows, const int *cols, double *ptr) const override {
        for (int k = 0; k < N; k++)
            for (int j = 0; j < M; j++)
                ptr[j + static_cast<std::size_t>(M) * k] = entry(rows[j], cols[k]);
    }
};

static double relative_error(const std::vector<double> &a, const std::vector<double> &ref) {
    double num = 0, den = 0;
    for (std::size_t i = 0; i < a.size(); i++) {
        num += (a[i] - ref[i]) * (a[i] - ref[i]);
        den += ref[i] * ref[i];
    }
    return std::sqrt(num / den);
}

int main(int argc, char *argv[]) {
    MPI_Init(&argc, &argv);
    int size = 1, rank = 0;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
    const int n      = argc > 1 ? std::atoi(argv[1]) : 20000;
    const int device = rank; // one process per GPU

    // geometry, clustering, compression: Htool on the host, exactly as in its own examples
    std::vector<double> points(3 * static_cast<std::size_t>(n));
    htool::create_sphere(n, points.data());
    htool::ClusterTreeBuilder<double> cluster_builder;
    htool::Cluster<double> cluster = cluster_builder.create_cluster_tree(n, 3, points.data(), 2, size);
    LaplaceKernel kernel(points);
    htool::HMatrixTreeBuilder<double, double> hmatrix_builder(1e-4, 10., 'N', 'N');
    htool::HMatrix<double, double> strip = hmatrix_builder.build(kernel, cluster, cluster, rank, rank); // this rank's block row (utility.hpp:56)

    std::vector<double> x(n, 1.), y_cpu(n, 0.), y_gpu(n, 0.);

    // (2) Htool's DistributedOperator: reference adapter vs GPU twin, same call
    htool::RestrictedGlobalToLocalHMatrix<double, double> cpu_operator(strip, strip.get_target_cluster(), strip.get_source_cluster(), false, false);
    htool_b200::RestrictedGlobalToLocalHMatrix<double, double> gpu_operator(strip, strip.get_target_cluster(), strip.get_source_cluster(), false, false, device);
    htool::CustomApproximationBuilder<double> cpu_builder(cluster, cluster, MPI_COMM_WORLD, cpu_operator);
    htool::CustomApproximationBuilder<double> gpu_builder(cluster, cluster, MPI_COMM_WORLD, gpu_operator);
    htool::add_distributed_operator_vector_product_global_to_global('N', 1., cpu_builder.distributed_operator, x.data(), 0., y_cpu.data(), static_cast<double *>(nullptr));
    htool::add_distributed_operator_vector_product_global_to_global('N', 1., gpu_builder.distributed_operator, x.data(), 0., y_gpu.data(), static_cast<double *>(nullptr));
    if (rank == 0)
        std::cout << "DistributedOperator, GPU twin vs reference adapter: " << relative_error(y_gpu, y_cpu) << "\n";

    // (3) one process per GPU, exchange over NVLink, same function names
    htool::PartitionFromCluster<double, double> partition(cluster);
    htool_b200::DeviceDistributedOperator<double> device_operator(strip, partition, MPI_COMM_WORLD, device);
    std::vector<double> y_dev(n, 0.);
    htool_b200::add_distributed_operator_vector_product_global_to_global('N', 1., device_operator, x.data(), 0., y_dev.data());
    if (rank == 0)
        std::cout << "DeviceDistributedOperator vs reference adapter:      " << relative_error(y_dev, y_cpu) << "\n";

    // the Krylov vectors of a solve are reused for every product: page-lock them once (zero-copy products afterwards)
    const int n_local = partition.get_size_of_partition(rank), offset = partition.get_offset_of_partition(rank);
    std::vector<double> rhs_global(n), rhs_local(n_local), solution_local(n_local, 0.);
    partition.global_to_partition_numbering(y_cpu.data(), rhs_global.data()); // rhs = A * ones: the solution is the vector of ones
    std::copy_n(rhs_global.begin() + offset, n_local, rhs_local.begin());
    htool_b200::PinnedHostBuffer<double> pin_rhs(rhs_local.data(), rhs_local.size()), pin_solution(solution_local.data(), solution_local.size());
    const htb_gmres_result result = device_operator.solve(rhs_local.data(), solution_local.data());
    if (rank == 0)
        std::cout << "device-resident GMRES: " << result.iterations << " iterations, relative residual " << result.true_relative_residual << ", x[0] = " << solution_local[0] << "\n";

    // (1) single process, user numbering (meaningful when the strip is the whole operator)
    if (size == 1) {
        htool_b200::DeviceHMatrix<double, double> device_hmatrix(strip, device);
        std::vector<double> y_user(n, 0.), y_ref(n, 0.);
        htool_b200::add_hmatrix_vector_product(exec_compat::par, 'N', 1., device_hmatrix, x.data(), 0., y_user.data());
        htool::add_hmatrix_vector_product(exec_compat::par, 'N', 1., strip, x.data(), 0., y_ref.data());
        std::cout << "add_hmatrix_vector_product (user numbering):         " << relative_error(y_user, y_ref) << "\n";
    }

    // (4) the WHOLE leaf assembly on the GPU: the same builder call, with generators that compute nothing plugged into the
    // builder's own hooks (tree_builder.hpp:251,258) — the host builds the block cluster tree, the device compresses the
    // admissible blocks with the reference's sympartialACA at the builder's epsilon and generates the dense leaves.
    {
        htool::HMatrixTreeBuilder<double, double> tree_builder(1e-4, 10., 'N', 'N');
        auto dense   = std::make_shared<htool_b200::DeviceDenseBlocks<double>>();
        auto lowrank = std::make_shared<htool_b200::DeviceLowRankBlocks<double>>();
        tree_builder.set_dense_blocks_generator(dense);
        tree_builder.set_low_rank_generator(std::static_pointer_cast<htool::VirtualInternalLowRankGenerator<double>>(lowrank));
        htool::HMatrix<double, double> tree = tree_builder.build(kernel, cluster, cluster, rank, rank); // no coefficient is computed
        htool_b200::BuiltinKernel builtin{HTB_KERNEL_LAPLACE_REG, 0., points.data(), points.data()};     // 1 / (1e-5 + 4 pi r), user numbering
        htool_b200::DeviceHMatrix<double, double> assembled(tree, *dense, *lowrank, builtin, device);
        const int n_rows = strip.get_target_cluster().get_size();
        std::vector<double> xc(n, 1.), y_dev_asm(n_rows, 0.), y_host_asm(n_rows, 0.);
        assembled.internal_add_vector_product('N', 1., xc.data(), 0., y_dev_asm.data());
        htool::openmp_internal_add_hmatrix_vector_product('N', 1., strip, xc.data(), 0., y_host_asm.data());
        if (rank == 0)
            std::cout << "device-assembled H-matrix vs host-assembled (strip product): " << relative_error(y_dev_asm, y_host_asm) << "\n";
        // ... and the distributed operator over the device-assembled strips (every rank compressed its own block row)
        htool_b200::DeviceDistributedOperator<double> assembled_operator(std::move(assembled), partition, MPI_COMM_WORLD);
        std::vector<double> y_asm(n, 0.);
        htool_b200::add_distributed_operator_vector_product_global_to_global('N', 1., assembled_operator, x.data(), 0., y_asm.data());
        if (rank == 0)
            std::cout << "DeviceDistributedOperator over device-assembled strips: " << relative_error(y_asm, y_cpu) << "\n";
    }
    MPI_Finalize();
    return 0;
}
