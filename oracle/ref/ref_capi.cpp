// oracle/ref/ref_capi.cpp — C entry points of oracle/_ref/libhtool_ref.so (TEST INFRASTRUCTURE, see
// ref_harness.hpp). Loaded with ctypes by tests/, tools/make_golden.py and bench.py (workload assembly by
// the reference + cpu_baseline / --impl reference timing). Never loaded by the product.
#include "ref_harness.hpp"

using htb_ref::CaseBase;

extern "C" {

void *ref_case_create(const ref_case_spec *spec) {
    try {
        return htb_ref::make_case(*spec);
    } catch (...) {
        return nullptr;
    }
}
void ref_case_destroy(void *h) { delete static_cast<CaseBase *>(h); }

const htb_hmatrix_desc *ref_case_desc(void *h) { return static_cast<CaseBase *>(h)->desc(); }
void ref_case_info(void *h, double *out, int n) { static_cast<CaseBase *>(h)->info(out, n); }
void ref_case_permutation(void *h, int side, int32_t *out) { static_cast<CaseBase *>(h)->permutation(side, out); }
void ref_case_points(void *h, int side, double *out) { static_cast<CaseBase *>(h)->points(side, out); }

// variant: 0 openmp_internal_*, 1 sequential_internal_*, 2 user-numbering add_hmatrix_vector_product(par),
//          3 through htool::LocalToLocalHMatrix, 4 through htool::RestrictedGlobalToLocalHMatrix
void ref_case_vector_product(void *h, int variant, char trans, const void *alpha, const void *in, const void *beta, void *out) {
    static_cast<CaseBase *>(h)->vector_product(variant, trans, alpha, in, beta, out);
}
void ref_case_matrix_product_row_major(void *h, int variant, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu) {
    static_cast<CaseBase *>(h)->matrix_product_row_major(variant, trans, alpha, in, beta, out, mu);
}
void ref_case_matrix_product_user(void *h, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu) {
    static_cast<CaseBase *>(h)->matrix_product_user(trans, alpha, in, beta, out, mu);
}
void ref_case_dense_product(void *h, char trans, const void *in, void *out) { static_cast<CaseBase *>(h)->dense_product(trans, in, out); }

void *ref_case_hmatrix(void *h) { return static_cast<CaseBase *>(h)->hmatrix_ptr(); }
void *ref_case_target_cluster(void *h) { return static_cast<CaseBase *>(h)->target_cluster_ptr(); }
void *ref_case_source_cluster(void *h) { return static_cast<CaseBase *>(h)->source_cluster_ptr(); }

void ref_set_num_threads(int n) { omp_set_num_threads(n); }
int ref_get_max_threads() { return omp_get_max_threads(); }
void ref_set_log_level(int level) { htool::Logger::get_instance().set_current_log_level(static_cast<unsigned int>(level)); }
}
