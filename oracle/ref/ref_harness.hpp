// oracle/ref/ref_harness.hpp — TEST INFRASTRUCTURE (oracle/): the UNMODIFIED reference, compiled from
// /root/reference/include where it lies, behind a small C API so that tests/, bench.py's cpu_baseline /
// --impl reference leg and tools/make_golden.py can (a) let the reference do what the north_star leaves
// to it — clustering, block tree, ACA/SVD compression on the host — and (b) run the reference's own CPU
// products on the very HMatrix object the GPU leaf store is flattened from.
// Nothing here is on the product path; the product (libhtool_b200.so) never links or loads this.
#ifndef HTB_ORACLE_REF_HARNESS_HPP
#define HTB_ORACLE_REF_HARNESS_HPP

#include <htool/htool.hpp>
#include <htool/testing/geometry.hpp>

#include <htool_b200/flatten.hpp>

#include <array>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <memory>
#include <omp.h>
#include <string>
#include <vector>

extern "C" {
typedef struct ref_case_spec {
    int32_t dtype;           // 0 double, 1 complex<double>
    int32_t kernel;          // 0: 1/(4 pi r) [generator_test.hpp:155-161]   1: 1/(1e-5+4 pi r) [:180-187]
                             // 2: (1+i)/(1e-5+4 pi r) [:189-196]            3: Hermitian (1 + sign(x_t-x_s) i)/(1e-5+4 pi r) [:198-205]
                             // 4: Helmholtz exp(ikr)/(4 pi r), finite diagonal (no reference generator)
                             // 5: (1+i)/(4 pi r) [:163-170]
    int32_t geometry_target; // 0: unit sphere surface (create_sphere normalised)  1: ball (create_sphere, geometry.hpp:46-61)  2: disk at z (create_disk, :41-44)
    int32_t geometry_source;
    int32_t n_target;
    int32_t n_source;
    int32_t same_cluster; // 1: source points/cluster tree are the target's (square operator; required for symmetry != 'N')
    int32_t min_depth;    // set_minimal_{target,source}_depth (tree_builder.hpp:256-257); 0 = default
    int32_t leaf_size;    // ClusterTreeBuilder::set_maximal_leaf_size; <=0 = default (10)
    int32_t n_partitions; // size_of_partition of the cluster trees (>= 1)
    int32_t partition_rank; // -1: whole operator. r >= 0: row strip of rank r, build(gen,target,source,r,r) (distributed_operator/utility.hpp:56)
    int32_t local_block;  // 1 (needs partition_rank >= 0): only the diagonal block, as DefaultLocalApproximationBuilder (utility.hpp:80)
    int32_t compressor;   // 0: default sympartialACA (tree_builder.hpp:385)  1: SVD  2: fullACA  3: partialACA
    int32_t symmetry;     // 'N' | 'S' | 'H'
    int32_t uplo;         // 'N' | 'L' | 'U'
    int32_t reserved;
    double z_target;
    double z_source;
    double epsilon;
    double eta;
    double wavenumber;
} ref_case_spec;
}

namespace htb_ref {

using complexd = std::complex<double>;

inline std::vector<double> make_points(int geometry, int n, double z) {
    std::vector<double> p(3 * static_cast<size_t>(n));
    if (geometry == 2) {
        htool::create_disk(3, z, n, p.data());
    } else {
        htool::create_sphere(n, p.data());
        if (geometry == 0) {
            for (int i = 0; i < n; i++) {
                double *q  = p.data() + 3 * static_cast<size_t>(i);
                double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
                if (nrm > 0) {
                    q[0] /= nrm;
                    q[1] /= nrm;
                    q[2] /= nrm;
                }
                q[2] += z;
            }
        } else if (z != 0) {
            for (int i = 0; i < n; i++)
                p[3 * static_cast<size_t>(i) + 2] += z;
        }
    }
    return p;
}

// The analytic kernels of include/htool/testing/generator_test.hpp, same formulas, as a VirtualGenerator
// in user numbering (interfaces/virtual_generator.hpp:21-32). Written with plain loops because the
// fixture's std::inner_product + virtual get_coef per coefficient is too slow for N = 1e6.
template <typename T>
class KernelGenerator final : public htool::VirtualGenerator<T> {
    const double *m_t;
    const double *m_s;
    int m_kernel;
    double m_k;

  public:
    KernelGenerator(const std::vector<double> &t, const std::vector<double> &s, int kernel, double wavenumber) : m_t(t.data()), m_s(s.data()), m_kernel(kernel), m_k(wavenumber) {}

    inline T coef(int i, int j) const {
        const double *a = m_t + 3 * static_cast<size_t>(i);
        const double *b = m_s + 3 * static_cast<size_t>(j);
        double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        return value(r, a[0] - b[0]);
    }
    void copy_submatrix(int M, int N, const int *rows, const int *cols, T *ptr) const override {
        for (int j = 0; j < N; j++)
            for (int i = 0; i < M; i++)
                ptr[i + static_cast<size_t>(M) * j] = coef(rows[i], cols[j]);
    }

  private:
    inline T value(double r, double dx0) const;
};

template <>
inline double KernelGenerator<double>::value(double r, double) const {
    if (m_kernel == 0)
        return 1. / (4 * M_PI * r);
    return 1. / (1e-5 + 4 * M_PI * r);
}
template <>
inline complexd KernelGenerator<complexd>::value(double r, double dx0) const {
    switch (m_kernel) {
    case 5:
        return (1. + complexd(0, 1)) / (4 * M_PI * r);
    case 3: {
        double s = dx0 > 0 ? 1. : (dx0 < 0 ? -1. : 0.);
        return (1. + s * complexd(0, 1)) / (1e-5 + 4 * M_PI * r);
    }
    case 4: {
        if (r < 1e-12)
            return complexd(1. / (4 * M_PI * 1e-3), m_k / (4 * M_PI)); // finite diagonal
        return std::exp(complexd(0, m_k * r)) / (4 * M_PI * r);
    }
    default:
        return (1. + complexd(0, 1)) / (1e-5 + 4 * M_PI * r);
    }
}

struct CaseBase {
    ref_case_spec spec{};
    double build_seconds{0}, cluster_seconds{0};
    virtual ~CaseBase() = default;
    virtual const htb_hmatrix_desc *desc()                                                                                             = 0;
    virtual void info(double *out, int n)                                                                                              = 0;
    virtual void permutation(int side, int32_t *out)                                                                                   = 0;
    virtual void points(int side, double *out)                                                                                         = 0;
    virtual void vector_product(int variant, char trans, const void *alpha, const void *in, const void *beta, void *out)               = 0;
    virtual void matrix_product_row_major(int variant, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu) = 0;
    virtual void matrix_product_user(char trans, const void *alpha, const void *in, const void *beta, void *out, int mu)               = 0;
    virtual void dense_product(char trans, const void *in, void *out)                                                                  = 0;
    virtual void *hmatrix_ptr()                                                                                                        = 0;
    virtual void *target_cluster_ptr()                                                                                                 = 0;
    virtual void *source_cluster_ptr()                                                                                                 = 0;
};

template <typename T>
struct Case final : CaseBase {
    using HMatrixType = htool::HMatrix<T, double>;
    std::vector<double> target_points, source_points_storage;
    const std::vector<double> *source_points{nullptr};
    std::unique_ptr<htool::Cluster<double>> target_cluster, source_cluster_storage;
    const htool::Cluster<double> *source_cluster{nullptr};
    std::unique_ptr<KernelGenerator<T>> generator;
    std::unique_ptr<HMatrixType> hmatrix;
    htool_b200::FlatHMatrix flat;

    explicit Case(const ref_case_spec &s) {
        spec = s;
        auto t0 = std::chrono::steady_clock::now();
        target_points = make_points(s.geometry_target, s.n_target, s.z_target);
        htool::ClusterTreeBuilder<double> cluster_builder;
        if (s.leaf_size > 0)
            cluster_builder.set_maximal_leaf_size(s.leaf_size);
        int P          = s.n_partitions > 0 ? s.n_partitions : 1;
        target_cluster = std::make_unique<htool::Cluster<double>>(cluster_builder.create_cluster_tree(s.n_target, 3, target_points.data(), 2, P));
        if (s.same_cluster) {
            source_points  = &target_points;
            source_cluster = target_cluster.get();
        } else {
            source_points_storage  = make_points(s.geometry_source, s.n_source, s.z_source);
            source_points          = &source_points_storage;
            source_cluster_storage = std::make_unique<htool::Cluster<double>>(cluster_builder.create_cluster_tree(s.n_source, 3, source_points_storage.data(), 2, P));
            source_cluster         = source_cluster_storage.get();
        }
        auto t1         = std::chrono::steady_clock::now();
        cluster_seconds = std::chrono::duration<double>(t1 - t0).count();

        generator = std::make_unique<KernelGenerator<T>>(target_points, *source_points, s.kernel, s.wavenumber);

        std::shared_ptr<htool::VirtualInternalLowRankGenerator<T>> compressor;
        const int *tp = target_cluster->get_permutation().data();
        const int *sp = source_cluster->get_permutation().data();
        // the internal generator the compressors sample, in cluster numbering (virtual_generator.hpp:35-49)
        internal_generator = std::make_unique<htool::InternalGeneratorWithPermutation<T>>(*generator, tp, sp);
        if (s.compressor == 1)
            compressor = std::make_shared<htool::SVD<T>>(*internal_generator);
        else if (s.compressor == 2)
            compressor = std::make_shared<htool::fullACA<T>>(*internal_generator);
        else if (s.compressor == 3)
            compressor = std::make_shared<htool::partialACA<T>>(*internal_generator);

        std::unique_ptr<htool::HMatrixTreeBuilder<T, double>> tree_builder;
        if (compressor)
            tree_builder = std::make_unique<htool::HMatrixTreeBuilder<T, double>>(s.epsilon, s.eta, static_cast<char>(s.symmetry), static_cast<char>(s.uplo), -1, compressor);
        else
            tree_builder = std::make_unique<htool::HMatrixTreeBuilder<T, double>>(s.epsilon, s.eta, static_cast<char>(s.symmetry), static_cast<char>(s.uplo));
        if (s.min_depth > 0) {
            tree_builder->set_minimal_target_depth(s.min_depth);
            tree_builder->set_minimal_source_depth(s.min_depth);
        }
        if (s.partition_rank >= 0 && s.local_block) {
            hmatrix = std::make_unique<HMatrixType>(tree_builder->build(exec_compat::par, *internal_generator, target_cluster->get_cluster_on_partition(s.partition_rank), source_cluster->get_cluster_on_partition(s.partition_rank)));
        } else if (s.partition_rank >= 0) {
            hmatrix = std::make_unique<HMatrixType>(tree_builder->build(exec_compat::par, *internal_generator, *target_cluster, *source_cluster, s.partition_rank, s.partition_rank));
        } else {
            hmatrix = std::make_unique<HMatrixType>(tree_builder->build(exec_compat::par, *internal_generator, *target_cluster, *source_cluster));
        }
        auto t2       = std::chrono::steady_clock::now();
        build_seconds = std::chrono::duration<double>(t2 - t1).count();
        flat          = htool_b200::flatten(*hmatrix);
    }

    std::unique_ptr<htool::InternalGeneratorWithPermutation<T>> internal_generator;

    const htb_hmatrix_desc *desc() override {
        flat.desc.leaves = flat.leaves.data();
        return &flat.desc;
    }

    void info(double *out, int n) override {
        std::vector<double> v(24, 0.);
        int64_t dense = 0, lr = 0, twice = 0, coef = 0, coef_twice = 0;
        int rmin = 1 << 30, rmax = -1;
        for (const auto &l : flat.leaves) {
            int64_t c = l.rank < 0 ? int64_t(l.nb_rows) * l.nb_cols : int64_t(l.rank) * (l.nb_rows + l.nb_cols);
            coef += c;
            if (l.rank < 0)
                dense++;
            else {
                lr++;
                rmin = std::min(rmin, l.rank);
                rmax = std::max(rmax, l.rank);
            }
            if (l.flags & HTB_LEAF_APPLY_TRANSPOSED_TOO) {
                twice++;
                coef_twice += c;
            }
        }
        v[0]  = flat.desc.nb_rows;
        v[1]  = flat.desc.nb_cols;
        v[2]  = flat.desc.row_offset;
        v[3]  = flat.desc.col_offset;
        v[4]  = double(flat.leaves.size());
        v[5]  = double(dense);
        v[6]  = double(lr);
        v[7]  = double(twice);
        v[8]  = double(coef);
        v[9]  = double(coef_twice);
        v[10] = lr ? rmin : 0;
        v[11] = rmax;
        v[12] = flat.desc.symmetry_for_leaves;
        v[13] = flat.desc.uplo_for_leaves;
        v[14] = build_seconds;
        v[15] = cluster_seconds;
        v[16] = target_cluster->get_size();
        v[17] = source_cluster->get_size();
        v[18] = omp_get_max_threads();
        for (int i = 0; i < n && i < int(v.size()); i++)
            out[i] = v[i];
    }

    void permutation(int side, int32_t *out) override {
        const auto &c   = side == 0 ? hmatrix->get_target_cluster() : hmatrix->get_source_cluster();
        const auto &prm = c.get_permutation();
        for (int i = 0; i < c.get_size(); i++)
            out[i] = prm[c.get_offset() + i] - c.get_offset();
    }

    // the points of the root block's rows (side 0) / columns (side 1) in CLUSTER numbering, 3 doubles each: what
    // htb_generator_desc wants (row i of the block <-> user index permutation[offset + i])
    void points(int side, double *out) override {
        const auto &c       = side == 0 ? hmatrix->get_target_cluster() : hmatrix->get_source_cluster();
        const auto &prm     = c.get_permutation();
        const double *pts   = side == 0 ? target_points.data() : source_points->data();
        for (int i = 0; i < c.get_size(); i++)
            for (int d = 0; d < 3; d++)
                out[3 * static_cast<size_t>(i) + d] = pts[3 * static_cast<size_t>(prm[c.get_offset() + i]) + d];
    }

    void vector_product(int variant, char trans, const void *alpha, const void *in, const void *beta, void *out) override {
        T a = *static_cast<const T *>(alpha), b = *static_cast<const T *>(beta);
        const T *x = static_cast<const T *>(in);
        T *y       = static_cast<T *>(out);
        switch (variant) {
        case 0:
            htool::openmp_internal_add_hmatrix_vector_product(trans, a, *hmatrix, x, b, y);
            break;
        case 1:
            htool::sequential_internal_add_hmatrix_vector_product(trans, a, *hmatrix, x, b, y);
            break;
        case 2:
            htool::add_hmatrix_vector_product(exec_compat::par, trans, a, *hmatrix, x, b, y);
            break;
        case 3: {
            htool::LocalToLocalHMatrix<T, double> op(*hmatrix);
            static_cast<const htool::VirtualLocalToLocalOperator<T> &>(op).add_vector_product(trans, a, x, b, y);
            break;
        }
        case 4: {
            // in is GLOBAL (source root numbering) for 'N', out local; swapped otherwise (virtual_global_to_local_operator.hpp:11-15)
            htool::RestrictedGlobalToLocalHMatrix<T, double> op(*hmatrix, hmatrix->get_target_cluster(), hmatrix->get_source_cluster(), false, false);
            static_cast<const htool::VirtualGlobalToLocalOperator<T> &>(op).add_vector_product(trans, a, x, b, y);
            break;
        }
        default:
            break;
        }
    }

    void matrix_product_row_major(int variant, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu) override {
        T a = *static_cast<const T *>(alpha), b = *static_cast<const T *>(beta);
        const T *x = static_cast<const T *>(in);
        T *y       = static_cast<T *>(out);
        switch (variant) {
        case 0:
            htool::openmp_internal_add_hmatrix_matrix_product_row_major(trans, 'N', a, *hmatrix, x, b, y, mu);
            break;
        case 1:
            htool::sequential_internal_add_hmatrix_matrix_product_row_major(trans, 'N', a, *hmatrix, x, b, y, mu);
            break;
        case 3: {
            htool::LocalToLocalHMatrix<T, double> op(*hmatrix);
            static_cast<const htool::VirtualLocalToLocalOperator<T> &>(op).add_matrix_product_row_major(trans, a, x, b, y, mu);
            break;
        }
        case 4: {
            htool::RestrictedGlobalToLocalHMatrix<T, double> op(*hmatrix, hmatrix->get_target_cluster(), hmatrix->get_source_cluster(), false, false);
            static_cast<const htool::VirtualGlobalToLocalOperator<T> &>(op).add_matrix_product_row_major(trans, a, x, b, y, mu);
            break;
        }
        default:
            break;
        }
    }

    void matrix_product_user(char trans, const void *alpha, const void *in, const void *beta, void *out, int mu) override {
        T a = *static_cast<const T *>(alpha), b = *static_cast<const T *>(beta);
        int ni = trans == 'N' ? hmatrix->nb_cols() : hmatrix->nb_rows();
        int no = trans == 'N' ? hmatrix->nb_rows() : hmatrix->nb_cols();
        htool::Matrix<T> B(ni, mu), C(no, mu);
        std::copy_n(static_cast<const T *>(in), size_t(ni) * mu, B.data());
        std::copy_n(static_cast<const T *>(out), size_t(no) * mu, C.data());
        htool::add_hmatrix_matrix_product(exec_compat::par, trans, 'N', a, *hmatrix, B, b, C);
        std::copy_n(C.data(), size_t(no) * mu, static_cast<T *>(out));
    }

    // Dense product of the same analytic kernel in CLUSTER numbering of the root block: what the
    // reference's tests compare against (test_hmatrix_matrix_product.hpp:127-156), O(m n).
    void dense_product(char trans, const void *in, void *out) override {
        const auto &tc = hmatrix->get_target_cluster();
        const auto &sc = hmatrix->get_source_cluster();
        const int m = tc.get_size(), n = sc.get_size();
        const int *tp = tc.get_permutation().data() + tc.get_offset();
        const int *sp = sc.get_permutation().data() + sc.get_offset();
        const T *x    = static_cast<const T *>(in);
        T *y          = static_cast<T *>(out);
        char sym      = static_cast<char>(spec.symmetry);
        (void)sym;
        if (trans == 'N') {
#pragma omp parallel for
            for (int i = 0; i < m; i++) {
                T acc = 0;
                for (int j = 0; j < n; j++)
                    acc += generator->coef(tp[i], sp[j]) * x[j];
                y[i] = acc;
            }
        } else {
#pragma omp parallel for
            for (int j = 0; j < n; j++) {
                T acc = 0;
                for (int i = 0; i < m; i++) {
                    T c = generator->coef(tp[i], sp[j]);
                    if (trans == 'C')
                        c = htool::conj_if_complex<T>(c);
                    acc += c * x[i];
                }
                y[j] = acc;
            }
        }
    }

    void *hmatrix_ptr() override { return hmatrix.get(); }
    void *target_cluster_ptr() override { return target_cluster.get(); }
    void *source_cluster_ptr() override { return const_cast<htool::Cluster<double> *>(source_cluster); }
};

inline CaseBase *make_case(const ref_case_spec &s) {
    if (s.dtype == 0)
        return new Case<double>(s);
    return new Case<complexd>(s);
}

} // namespace htb_ref
#endif
