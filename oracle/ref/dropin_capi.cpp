// oracle/ref/dropin_capi.cpp — TEST INFRASTRUCTURE (oracle/_ref/libhtool_dropin.so): the reference's own
// operator and DistributedOperator code (unmodified headers under /root/reference/include) driving, side by
// side, (a) the reference's CPU adapters and (b) the GPU twins of htool_b200/cpp/htool_b200/operators.hpp
// on the SAME HMatrix object. It is the drop-in proof of SURVEY.md 8b/8c: every call below is the
// reference's call with only the operator object swapped. Linked against libhtool_b200.so; needs a GPU at
// run time (tests/test_gpu_dropin.py, -m gpu). MPI is the single-rank stub of oracle/ref/mpi_stub.
#include "ref_harness.hpp"

// htool.hpp only pulls the global-to-global linalg (distributed_operator/linalg.hpp); the local-to-local and
// sub-product entry points HPDDM / GenEO use are separate headers
#include <htool/distributed_operator/linalg/add_distributed_operator_matrix_product_local_to_local.hpp>
#include <htool/distributed_operator/linalg/add_distributed_operator_matrix_product_row_major_local_to_local.hpp>
#include <htool/distributed_operator/linalg/add_distributed_operator_vector_product_local_to_local.hpp>
#include <htool/distributed_operator/linalg/add_distributed_operator_vector_sub_product_global_to_local.hpp>

#include <htool_b200/distributed.hpp>
#include <htool_b200/operators.hpp>

#include <cstdio>
#include <cstdlib>
#include <random>

namespace {

using htb_ref::complexd;

template <typename T>
T rnd_scalar(std::mt19937 &g);
template <>
double rnd_scalar<double>(std::mt19937 &g) { return std::uniform_real_distribution<double>(-1., 1.)(g); }
template <>
complexd rnd_scalar<complexd>(std::mt19937 &g) {
    std::uniform_real_distribution<double> d(-1., 1.);
    double a = d(g), b = d(g);
    return complexd(a, b);
}
template <typename T>
std::vector<T> rnd_vector(std::mt19937 &g, size_t n) {
    std::vector<T> v(n);
    for (auto &x : v)
        x = rnd_scalar<T>(g);
    return v;
}
template <typename T>
double rel_err(const std::vector<T> &a, const std::vector<T> &b) {
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); i++) {
        num += std::norm(a[i] - b[i]);
        den += std::norm(b[i]);
    }
    return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}

std::vector<char> valid_trans(char sym, bool is_complex) {
    std::vector<char> t{'N'};
    if (sym != 'H')
        t.push_back('T');
    if (is_complex && sym != 'S')
        t.push_back('C');
    return t;
}

// results[k]: worst relative l2 difference GPU-twin vs reference for check group k
enum { G2L_VECTOR = 0,
       G2L_ROW_MAJOR,
       G2L_SUB_PRODUCT,
       L2L_VECTOR,
       L2L_ROW_MAJOR,
       L2L_SUB_PRODUCT,
       DIST_VECTOR_G2G,
       DIST_VECTOR_L2L,
       DIST_MATRIX_G2G,
       DIST_MATRIX_L2L,
       DIST_ROW_MAJOR_L2L,
       DIST_SUB_PRODUCT,
       FREE_VECTOR_USER,
       FREE_MATRIX_USER,
       LOGGED_UNSUPPORTED,
       DEVICE_DIST,
       GENERATED_DENSE,
       DEVICE_ASSEMBLY,
       N_GROUPS };

struct CountingWriter : htool::IObjectWriter {
    int errors = 0;
    void set_log_level(htool::LogLevel) override {}
    void write(htool::LogLevel level, const std::string &) override {
        if (level <= htool::LogLevel::ERROR)
            errors++;
    }
};

template <typename T>
int run(htb_ref::Case<T> &c, double *results, int n_results) {
    using namespace htool;
    const bool is_complex = !std::is_same<T, double>::value;
    std::vector<double> worst(N_GROUPS, -1.); // -1: group not run for this case
    const bool verbose = std::getenv("HTB_DROPIN_VERBOSE") != nullptr;
    auto upd = [&](int k, double e) {
        worst[k] = std::max(worst[k], e);
        if (verbose)
            std::fprintf(stderr, "[dropin] group %d err %.3e\n", k, e);
    };
    std::mt19937 gen(7);
    const HMatrix<T, double> &H = *c.hmatrix;
    const char sym              = H.get_symmetry_for_leaves();
    const int nr = H.get_target_cluster().get_size(), nc = H.get_source_cluster().get_size();
    const int NT = c.target_cluster->get_size(), NS = c.source_cluster->get_size();
    const bool whole = c.spec.partition_rank < 0;

    // ---- operator objects: reference vs twin, same constructor arguments -------------------------------------
    htool::RestrictedGlobalToLocalHMatrix<T, double> ref_g2l(H, H.get_target_cluster(), H.get_source_cluster(), false, false);
    htool_b200::RestrictedGlobalToLocalHMatrix<T, double> gpu_g2l(H, H.get_target_cluster(), H.get_source_cluster(), false, false);
    if (!gpu_g2l.get_device_hmatrix().is_valid())
        return 1;
    const VirtualGlobalToLocalOperator<T> &rg = ref_g2l, &gg = gpu_g2l;

    for (char trans : valid_trans(sym, is_complex)) {
        // g2l: 'N' in global / out local, else in local / out global (virtual_global_to_local_operator.hpp:11-15)
        const size_t ni = trans == 'N' ? NS : nr, no = trans == 'N' ? nr : NS;
        for (int mu : {1, 5}) {
            T alpha = rnd_scalar<T>(gen), beta = rnd_scalar<T>(gen);
            auto x = rnd_vector<T>(gen, ni * mu), y0 = rnd_vector<T>(gen, no * mu);
            auto yr = y0, yg = y0;
            if (mu == 1) {
                rg.add_vector_product(trans, alpha, x.data(), beta, yr.data());
                gg.add_vector_product(trans, alpha, x.data(), beta, yg.data());
                upd(G2L_VECTOR, rel_err(yg, yr));
            }
            yr = y0, yg = y0;
            rg.add_matrix_product_row_major(trans, alpha, x.data(), beta, yr.data(), mu);
            gg.add_matrix_product_row_major(trans, alpha, x.data(), beta, yg.data(), mu);
            upd(G2L_ROW_MAJOR, rel_err(yg, yr));
        }
    }
    {
        // sub-products: aligned with the source range, clipped on either side, disjoint (coarse_operator_builder.hpp:99)
        const int so = H.get_source_cluster().get_offset();
        const int mu = 3;
        const int cases[][2] = {{so, nc}, {so + nc / 4, nc / 2}, {std::max(0, so - 5), nc / 3 + 5}, {so + nc / 2, nc - nc / 2}};
        for (auto &oc : cases) {
            int offset = oc[0], size = oc[1];
            if (size <= 0)
                continue;
            auto x = rnd_vector<T>(gen, size_t(size) * mu), y0 = rnd_vector<T>(gen, size_t(nr) * mu);
            auto yr = y0, yg = y0;
            rg.add_sub_matrix_product_to_local(x.data(), yr.data(), mu, offset, size);
            gg.add_sub_matrix_product_to_local(x.data(), yg.data(), mu, offset, size);
            upd(G2L_SUB_PRODUCT, rel_err(yg, yr));
        }
    }

    // ---- local-to-local adapters (in/out both local to the block) ----------------------------------------------
    {
        htool::LocalToLocalHMatrix<T, double> ref_l2l(H);
        htool_b200::LocalToLocalHMatrix<T, double> gpu_l2l(H);
        const VirtualLocalToLocalOperator<T> &rl = ref_l2l, &gl = gpu_l2l;
        for (char trans : valid_trans(sym, is_complex)) {
            const size_t ni = trans == 'N' ? nc : nr, no = trans == 'N' ? nr : nc;
            T alpha = rnd_scalar<T>(gen), beta = rnd_scalar<T>(gen);
            auto x = rnd_vector<T>(gen, ni), y0 = rnd_vector<T>(gen, no);
            auto yr = y0, yg = y0;
            rl.add_vector_product(trans, alpha, x.data(), beta, yr.data());
            gl.add_vector_product(trans, alpha, x.data(), beta, yg.data());
            upd(L2L_VECTOR, rel_err(yg, yr));
            const int mu = 4;
            auto X = rnd_vector<T>(gen, ni * mu), Y0 = rnd_vector<T>(gen, no * mu);
            auto Yr = Y0, Yg = Y0;
            rl.add_matrix_product_row_major(trans, alpha, X.data(), beta, Yr.data(), mu);
            gl.add_matrix_product_row_major(trans, alpha, X.data(), beta, Yg.data(), mu);
            upd(L2L_ROW_MAJOR, rel_err(Yg, Yr));
        }
        const int so = H.get_source_cluster().get_offset();
        const int mu = 1; // mu = 1: the reference advances `in` by rows, not rows*mu (local_to_local_operators/hmatrix.hpp:45)
        const int cases[][2] = {{so, nc}, {so + nc / 3, nc / 3}, {so + nc / 2, nc - nc / 2}};
        for (auto &oc : cases) {
            auto x = rnd_vector<T>(gen, size_t(oc[1]) * mu), y0 = rnd_vector<T>(gen, size_t(nr) * mu);
            auto yr = y0, yg = y0;
            rl.add_sub_matrix_product_to_local(x.data(), yr.data(), mu, oc[0], oc[1]);
            gl.add_sub_matrix_product_to_local(x.data(), yg.data(), mu, oc[0], oc[1]);
            upd(L2L_SUB_PRODUCT, rel_err(yg, yr));
        }
    }

    // ---- DistributedOperator: the reference's linalg, unchanged, with either operator plugged in ------------------
    // (single-rank stub MPI: meaningful when the H-matrix is the whole operator)
    if (whole) {
        CustomApproximationBuilder<T> ref_builder(*c.target_cluster, *c.source_cluster, MPI_COMM_WORLD, rg);
        CustomApproximationBuilder<T> gpu_builder(*c.target_cluster, *c.source_cluster, MPI_COMM_WORLD, gg);
        const DistributedOperator<T> &RA = ref_builder.distributed_operator, &GA = gpu_builder.distributed_operator;
        for (char trans : valid_trans(sym, is_complex)) {
            const size_t ni = trans == 'N' ? NS : NT, no = trans == 'N' ? NT : NS;
            T alpha = rnd_scalar<T>(gen), beta = rnd_scalar<T>(gen);
            // vector, user numbering, global to global (use_distributed_operator.cpp:108)
            auto x = rnd_vector<T>(gen, ni), y0 = rnd_vector<T>(gen, no);
            auto yr = y0, yg = y0;
            add_distributed_operator_vector_product_global_to_global(trans, alpha, RA, x.data(), beta, yr.data(), static_cast<T *>(nullptr));
            add_distributed_operator_vector_product_global_to_global(trans, alpha, GA, x.data(), beta, yg.data(), static_cast<T *>(nullptr));
            upd(DIST_VECTOR_G2G, rel_err(yg, yr));
            // vector, local to local: the call HPDDMOperator::GMV makes per Krylov iteration (wrapper_hpddm.hpp:120)
            yr = y0, yg = y0;
            std::vector<T> work(NS + NT + ni + no);
            internal_add_distributed_operator_vector_product_local_to_local(trans, alpha, RA, x.data(), beta, yr.data(), work.data());
            internal_add_distributed_operator_vector_product_local_to_local(trans, alpha, GA, x.data(), beta, yg.data(), work.data());
            upd(DIST_VECTOR_L2L, rel_err(yg, yr));
            yr = y0, yg = y0;
            add_distributed_operator_vector_product_local_to_local(trans, alpha, RA, x.data(), beta, yr.data(), static_cast<T *>(nullptr));
            add_distributed_operator_vector_product_local_to_local(trans, alpha, GA, x.data(), beta, yg.data(), static_cast<T *>(nullptr));
            upd(DIST_VECTOR_L2L, rel_err(yg, yr));
            // matrices (column-major, mu columns)
            const int mu = 5;
            Matrix<T> X(ni, mu), Yr(no, mu), Yg(no, mu);
            auto xv = rnd_vector<T>(gen, ni * mu), yv = rnd_vector<T>(gen, no * mu);
            std::copy(xv.begin(), xv.end(), X.data());
            std::copy(yv.begin(), yv.end(), Yr.data());
            std::copy(yv.begin(), yv.end(), Yg.data());
            add_distributed_operator_matrix_product_global_to_global(trans, alpha, RA, X, beta, Yr, static_cast<T *>(nullptr));
            add_distributed_operator_matrix_product_global_to_global(trans, alpha, GA, X, beta, Yg, static_cast<T *>(nullptr));
            upd(DIST_MATRIX_G2G, rel_err(std::vector<T>(Yg.data(), Yg.data() + no * mu), std::vector<T>(Yr.data(), Yr.data() + no * mu)));
            std::copy(yv.begin(), yv.end(), Yr.data());
            std::copy(yv.begin(), yv.end(), Yg.data());
            add_distributed_operator_matrix_product_local_to_local(trans, alpha, RA, X, beta, Yr, static_cast<T *>(nullptr));
            add_distributed_operator_matrix_product_local_to_local(trans, alpha, GA, X, beta, Yg, static_cast<T *>(nullptr));
            upd(DIST_MATRIX_L2L, rel_err(std::vector<T>(Yg.data(), Yg.data() + no * mu), std::vector<T>(Yr.data(), Yr.data() + no * mu)));
            // row-major local to local: GMV with mu > 1 (wrapper_hpddm.hpp:124)
            Matrix<T> Xr(mu, ni), Yrr(mu, no), Ygr(mu, no); // a (mu x n) column-major matrix is n x mu row-major
            std::copy(xv.begin(), xv.end(), Xr.data());
            std::copy(yv.begin(), yv.end(), Yrr.data());
            std::copy(yv.begin(), yv.end(), Ygr.data());
            std::vector<T> workm((NS + NT + ni + no) * size_t(mu) * 2);
            internal_add_distributed_operator_matrix_product_row_major_local_to_local(trans, alpha, RA, Xr, beta, Yrr, workm.data());
            internal_add_distributed_operator_matrix_product_row_major_local_to_local(trans, alpha, GA, Xr, beta, Ygr, workm.data());
            upd(DIST_ROW_MAJOR_L2L, rel_err(std::vector<T>(Ygr.data(), Ygr.data() + no * mu), std::vector<T>(Yrr.data(), Yrr.data() + no * mu)));
        }
        {
            const int mu = 2, offset = NS / 3, size = NS / 2;
            auto x = rnd_vector<T>(gen, size_t(size) * mu), y0 = rnd_vector<T>(gen, size_t(NT) * mu);
            auto yr = y0, yg = y0;
            internal_add_distributed_operator_vector_sub_product_global_to_local(RA, x.data(), yr.data(), mu, offset, size);
            internal_add_distributed_operator_vector_sub_product_global_to_local(GA, x.data(), yg.data(), mu, offset, size);
            upd(DIST_SUB_PRODUCT, rel_err(yg, yr));
        }
        // ---- DeviceDistributedOperator (htool_b200/distributed.hpp): the NCCL-backed twin of the linalg above, same names,
        // against the reference's DistributedOperator (one rank here: the collectives degenerate, the call path does not)
        if (c.spec.same_cluster) {
            htool_b200::DeviceDistributedOperator<T, double> DA(H, RA.get_target_partition(), MPI_COMM_WORLD);
            if (DA.is_valid()) {
                for (char trans : valid_trans(sym, is_complex)) {
                    T alpha = rnd_scalar<T>(gen), beta = rnd_scalar<T>(gen);
                    auto x = rnd_vector<T>(gen, NS), y0 = rnd_vector<T>(gen, NT);
                    auto yr = y0, yg = y0;
                    std::vector<T> work(2 * size_t(NS + NT));
                    internal_add_distributed_operator_vector_product_local_to_local(trans, alpha, RA, x.data(), beta, yr.data(), work.data());
                    htool_b200::internal_add_distributed_operator_vector_product_local_to_local(trans, alpha, DA, x.data(), beta, yg.data());
                    upd(DEVICE_DIST, rel_err(yg, yr));
                    yr = y0, yg = y0;
                    internal_add_distributed_operator_vector_product_global_to_global(trans, alpha, RA, x.data(), beta, yr.data(), static_cast<T *>(nullptr));
                    htool_b200::internal_add_distributed_operator_vector_product_global_to_global(trans, alpha, DA, x.data(), beta, yg.data());
                    upd(DEVICE_DIST, rel_err(yg, yr));
                    yr = y0, yg = y0;
                    add_distributed_operator_vector_product_global_to_global(trans, alpha, RA, x.data(), beta, yr.data(), static_cast<T *>(nullptr));
                    htool_b200::add_distributed_operator_vector_product_global_to_global(trans, alpha, DA, x.data(), beta, yg.data());
                    upd(DEVICE_DIST, rel_err(yg, yr));
                    const int mu = 3;
                    Matrix<T> Xr(mu, NS), Yrr(mu, NT);
                    auto xv = rnd_vector<T>(gen, size_t(NS) * mu), yv = rnd_vector<T>(gen, size_t(NT) * mu);
                    std::copy(xv.begin(), xv.end(), Xr.data());
                    std::copy(yv.begin(), yv.end(), Yrr.data());
                    auto ygm = yv;
                    std::vector<T> workm(2 * size_t(NS + NT) * mu);
                    internal_add_distributed_operator_matrix_product_row_major_local_to_local(trans, alpha, RA, Xr, beta, Yrr, workm.data());
                    htool_b200::internal_add_distributed_operator_matrix_product_row_major_local_to_local(trans, alpha, DA, xv.data(), beta, ygm.data(), mu);
                    upd(DEVICE_DIST, rel_err(ygm, std::vector<T>(Yrr.data(), Yrr.data() + size_t(NT) * mu)));
                }
                // device-resident GMRES through the twin; the residual is measured with the REFERENCE's product
                if (NS == NT) {
                    htb_gmres_options gopt;
                    htb_gmres_default_options(&gopt);
                    gopt.tolerance = 1e-8;
                    auto b         = rnd_vector<T>(gen, NT);
                    std::vector<T> xs(NS, T(0)), r = b, work(2 * size_t(NS + NT));
                    const htb_gmres_result gr = DA.solve(b.data(), xs.data(), &gopt);
                    if (gr.converged) {
                        internal_add_distributed_operator_vector_product_local_to_local('N', T(-1), RA, xs.data(), T(1), r.data(), work.data());
                        double rn = 0, bn = 0;
                        for (int i = 0; i < NT; i++) {
                            rn += std::norm(r[i]);
                            bn += std::norm(b[i]);
                        }
                        const double rel = std::sqrt(rn / bn);
                        upd(DEVICE_DIST, rel < 1e-7 ? 0. : rel);
                    }
                }
            } else {
                upd(DEVICE_DIST, 1.);
            }
        }
        // ---- free functions, user numbering (use_hmatrix.cpp:107) -------------------------------------------------
        // NOTE the reference's user-numbering front ends permute `in` with the SOURCE cluster and `out` with the TARGET
        // cluster whatever trans is (add_hmatrix_vector_product.hpp:178-179,196): for trans != 'N' they are only
        // meaningful (and only stay inside the caller's buffers) when both clusters are the same object, so that is
        // the only case compared here. The twin permutes the input of op(H) with the cluster of its input.
        htool_b200::DeviceHMatrix<T, double> DH(H);
        for (char trans : valid_trans(sym, is_complex)) {
            if (trans != 'N' && !c.spec.same_cluster)
                continue;
            const size_t ni = trans == 'N' ? NS : NT, no = trans == 'N' ? NT : NS;
            T alpha = rnd_scalar<T>(gen), beta = rnd_scalar<T>(gen);
            auto x = rnd_vector<T>(gen, ni), y0 = rnd_vector<T>(gen, no);
            auto yr = y0, yg = y0;
            htool::add_hmatrix_vector_product(exec_compat::par, trans, alpha, H, x.data(), beta, yr.data());
            htool_b200::add_hmatrix_vector_product(exec_compat::par, trans, alpha, DH, x.data(), beta, yg.data());
            upd(FREE_VECTOR_USER, rel_err(yg, yr));
            const int mu = 3;
            Matrix<T> B(ni, mu), Cr(no, mu);
            auto xv = rnd_vector<T>(gen, ni * mu), yv = rnd_vector<T>(gen, no * mu);
            std::copy(xv.begin(), xv.end(), B.data());
            std::copy(yv.begin(), yv.end(), Cr.data());
            auto cg = yv;
            htool::add_hmatrix_matrix_product(exec_compat::par, trans, 'N', alpha, H, B, beta, Cr);
            htool_b200::add_hmatrix_matrix_product(exec_compat::par, trans, 'N', alpha, DH, xv.data(), beta, cg.data(), mu);
            upd(FREE_MATRIX_USER, rel_err(cg, std::vector<T>(Cr.data(), Cr.data() + no * mu)));
        }
    }

    // ---- leaf assembly on the device: the SAME builder call with htool_b200::DeviceDenseBlocks as its dense-blocks generator
    // (HMatrixTreeBuilder::set_dense_blocks_generator, tree_builder.hpp:258): the dense leaves are never computed on the
    // host, the GPU generates them from the built-in kernel function; products must agree with the reference's on H -------
    if (c.spec.compressor == 0 && (c.spec.kernel >= 0 && c.spec.kernel <= 5)) {
        HMatrixTreeBuilder<T, double> builder(c.spec.epsilon, c.spec.eta, static_cast<char>(c.spec.symmetry), static_cast<char>(c.spec.uplo));
        if (c.spec.min_depth > 0) {
            builder.set_minimal_target_depth(c.spec.min_depth);
            builder.set_minimal_source_depth(c.spec.min_depth);
        }
        auto deferred = std::make_shared<htool_b200::DeviceDenseBlocks<T>>();
        builder.set_dense_blocks_generator(deferred);
        std::unique_ptr<HMatrix<T, double>> H2;
        if (c.spec.partition_rank >= 0 && c.spec.local_block)
            H2 = std::make_unique<HMatrix<T, double>>(builder.build(exec_compat::par, *c.internal_generator, c.target_cluster->get_cluster_on_partition(c.spec.partition_rank), c.source_cluster->get_cluster_on_partition(c.spec.partition_rank)));
        else if (c.spec.partition_rank >= 0)
            H2 = std::make_unique<HMatrix<T, double>>(builder.build(exec_compat::par, *c.internal_generator, *c.target_cluster, *c.source_cluster, c.spec.partition_rank, c.spec.partition_rank));
        else
            H2 = std::make_unique<HMatrix<T, double>>(builder.build(exec_compat::par, *c.internal_generator, *c.target_cluster, *c.source_cluster));
        htool_b200::BuiltinKernel bk;
        bk.kernel        = c.spec.kernel;
        bk.wavenumber    = c.spec.wavenumber;
        bk.target_points = c.target_points.data();
        bk.source_points = c.source_points->data();
        htool_b200::DeviceHMatrix<T, double> DG(*H2, *deferred, bk);
        if (!DG.is_valid()) { // (deferred->size() may be 0: two well separated geometries have no near field at all)
            upd(GENERATED_DENSE, 1.);
        } else {
            for (char trans : valid_trans(sym, is_complex)) {
                const size_t ni = trans == 'N' ? nc : nr, no = trans == 'N' ? nr : nc;
                T alpha = rnd_scalar<T>(gen), beta = rnd_scalar<T>(gen);
                auto x = rnd_vector<T>(gen, ni), y0 = rnd_vector<T>(gen, no);
                auto yr = y0, yg = y0;
                openmp_internal_add_hmatrix_vector_product(trans, alpha, H, x.data(), beta, yr.data());
                DG.internal_add_vector_product(trans, alpha, x.data(), beta, yg.data());
                upd(GENERATED_DENSE, rel_err(yg, yr));
            }
        }
    }

    // ---- the WHOLE leaf assembly on the device: the same builder call with DeviceDenseBlocks AND DeviceLowRankBlocks
    // (HMatrixTreeBuilder::set_low_rank_generator, tree_builder.hpp:251): the host builds the block cluster tree only, the
    // GPU compresses the admissible blocks with the reference's sympartialACA and generates the dense leaves. Same ranks as
    // the reference's leaves, products equal to the reference's on H (all the built-in kernel functions) ------------------
    if (c.spec.compressor == 0 && (c.spec.kernel >= 0 && c.spec.kernel <= 5)) {
        HMatrixTreeBuilder<T, double> builder(c.spec.epsilon, c.spec.eta, static_cast<char>(c.spec.symmetry), static_cast<char>(c.spec.uplo));
        if (c.spec.min_depth > 0) {
            builder.set_minimal_target_depth(c.spec.min_depth);
            builder.set_minimal_source_depth(c.spec.min_depth);
        }
        auto deferred = std::make_shared<htool_b200::DeviceDenseBlocks<T>>();
        auto lowrank  = std::make_shared<htool_b200::DeviceLowRankBlocks<T>>();
        builder.set_dense_blocks_generator(deferred);
        builder.set_low_rank_generator(std::static_pointer_cast<VirtualInternalLowRankGenerator<T>>(lowrank));
        std::unique_ptr<HMatrix<T, double>> H2;
        if (c.spec.partition_rank >= 0 && c.spec.local_block)
            H2 = std::make_unique<HMatrix<T, double>>(builder.build(exec_compat::par, *c.internal_generator, c.target_cluster->get_cluster_on_partition(c.spec.partition_rank), c.source_cluster->get_cluster_on_partition(c.spec.partition_rank)));
        else if (c.spec.partition_rank >= 0)
            H2 = std::make_unique<HMatrix<T, double>>(builder.build(exec_compat::par, *c.internal_generator, *c.target_cluster, *c.source_cluster, c.spec.partition_rank, c.spec.partition_rank));
        else
            H2 = std::make_unique<HMatrix<T, double>>(builder.build(exec_compat::par, *c.internal_generator, *c.target_cluster, *c.source_cluster));
        htool_b200::BuiltinKernel bk;
        bk.kernel        = c.spec.kernel;
        bk.wavenumber    = c.spec.wavenumber;
        bk.target_points = c.target_points.data();
        bk.source_points = c.source_points->data();
        htool_b200::DeviceHMatrix<T, double> DA(*H2, *deferred, *lowrank, bk);
        if (!DA.is_valid()) {
            upd(DEVICE_ASSEMBLY, 1.);
        } else {
            // ranks: leaf by leaf against the H-matrix the reference assembled (same tree, same order of get_leaves_from)
            auto ranks       = DA.leaf_ranks();
            auto ref_leaves  = htool::get_leaves_from(H).first;
            std::size_t k    = 0;
            bool ranks_equal = true;
            for (const auto *leaf : ref_leaves) {
                if (!leaf->is_dense() && !leaf->is_low_rank())
                    continue;
                const int r = leaf->is_dense() ? -1 : leaf->get_low_rank_data()->rank_of();
                ranks_equal = ranks_equal && k < ranks.size() && ranks[k] == r;
                k++;
            }
            upd(DEVICE_ASSEMBLY, ranks_equal && k == ranks.size() ? 0. : 1.);
            for (char trans : valid_trans(sym, is_complex)) {
                const size_t ni = trans == 'N' ? nc : nr, no = trans == 'N' ? nr : nc;
                T alpha = rnd_scalar<T>(gen), beta = rnd_scalar<T>(gen);
                auto x = rnd_vector<T>(gen, ni), y0 = rnd_vector<T>(gen, no);
                auto yr = y0, yg = y0;
                openmp_internal_add_hmatrix_vector_product(trans, alpha, H, x.data(), beta, yr.data());
                DA.internal_add_vector_product(trans, alpha, x.data(), beta, yg.data());
                upd(DEVICE_ASSEMBLY, rel_err(yg, yr));
            }
        }
    }

    // ---- error convention: an unsupported trans is logged through htool::Logger, nothing is thrown ------------------
    if (sym == 'S' || sym == 'H') {
        auto writer = std::make_shared<CountingWriter>();
        Logger::get_instance().set_current_writer(writer);
        char bad = sym == 'S' ? 'C' : 'T';
        std::vector<T> x(std::max(NS, nr), T(1)), y(std::max(NS, nr), T(0));
        gg.add_vector_product(bad, T(1), x.data(), T(0), y.data());
        worst[LOGGED_UNSUPPORTED] = writer->errors > 0 ? 0. : 1.;
        Logger::get_instance().set_current_writer(std::make_shared<StandartOutputWriter>());
    }

    for (int i = 0; i < n_results && i < N_GROUPS; i++)
        results[i] = worst[i];
    return 0;
}

} // namespace

extern "C" {

int dropin_n_groups() { return N_GROUPS; }

// Builds the case with the reference (same spec as ref_case_create) and runs every check group.
// Returns 0 on success, 1 when the device leaf store could not be created, 2 on exception.
int dropin_run(const ref_case_spec *spec, double *results, int n_results) {
    try {
        std::unique_ptr<htb_ref::CaseBase> base(htb_ref::make_case(*spec));
        if (spec->dtype == 0)
            return run(*static_cast<htb_ref::Case<double> *>(base.get()), results, n_results);
        return run(*static_cast<htb_ref::Case<htb_ref::complexd> *>(base.get()), results, n_results);
    } catch (...) {
        return 2;
    }
}
}
