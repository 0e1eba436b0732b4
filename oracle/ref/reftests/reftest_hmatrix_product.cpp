// oracle/ref/reftests/reftest_hmatrix_product.cpp — TEST INFRASTRUCTURE. The reference's OWN test program
//   tests/functional_tests/hmatrix/hmatrix_product/test_hmatrix_product_{double,complex_double}.cpp
//   (-> test_hmatrix_product.hpp -> test_hmatrix_matrix_product.hpp:36-184 plain, :187-300 symmetric, :303-412 Hermitian)
// compiled UNMODIFIED where it lies under /root/reference, with the H-matrix x vector / x matrix products it checks
// against dense products routed to the GPU:
//   1. every reference header is included first (include guards set, the reference's functions keep their names);
//   2. wrappers with a b200_ prefix are defined: given an htool::HMatrix they upload it (htool_b200::DeviceHMatrix, one
//      upload per call: the test's H-matrices are short-lived locals) and call the C ABI; every other argument list is
//      forwarded to the reference's function of the same name (H x H, H x lrmat ... are not on the hot path);
//   3. #defines make the CALL SITES of the test bodies that follow use the wrappers;
//   4. the reference's test main() follows.
// Exit code 0 = the reference's own tolerances (error vs the dense product < epsilon in {1e-6, 1e-10}) hold on the GPU.
#include <htool/htool.hpp>
#include <htool/testing/generate_test_case.hpp>
#include <htool/testing/generator_input.hpp>
#include <htool/testing/generator_test.hpp>

#include <htool_b200/device_hmatrix.hpp>

#include <utility>

namespace htool {

// ---- vector products (cluster numbering): add_hmatrix_vector_product.hpp:17-54 (dispatch), :107-170 (openmp) ----------
template <typename T, typename U>
void b200_internal_add_hmatrix_vector_product(char trans, T alpha, const HMatrix<T, U> &A, const T *in, T beta, T *out) {
    htool_b200::DeviceHMatrix<T, U>(A).internal_add_vector_product(trans, alpha, in, beta, out);
}
template <typename... Args>
void b200_internal_add_hmatrix_vector_product(Args &&...args) { internal_add_hmatrix_vector_product(std::forward<Args>(args)...); }

template <typename T, typename U>
void b200_openmp_internal_add_hmatrix_vector_product(char trans, T alpha, const HMatrix<T, U> &A, const T *in, T beta, T *out) {
    htool_b200::DeviceHMatrix<T, U>(A).internal_add_vector_product(trans, alpha, in, beta, out);
}

// ---- row-major multi-RHS products: add_hmatrix_matrix_product_row_major.hpp:58-109 (sequential), :112-178 (openmp) -------
template <typename T, typename U>
void b200_openmp_internal_add_hmatrix_matrix_product_row_major(char transa, char transb, T alpha, const HMatrix<T, U> &A, const T *in, T beta, T *out, int mu) {
    htool_b200::internal_add_hmatrix_matrix_product_row_major(transa, transb, alpha, htool_b200::DeviceHMatrix<T, U>(A), in, beta, out, mu);
}
template <typename T, typename U>
void b200_sequential_internal_add_hmatrix_matrix_product_row_major(char transa, char transb, T alpha, const HMatrix<T, U> &A, const T *in, T beta, T *out, int mu) {
    htool_b200::internal_add_hmatrix_matrix_product_row_major(transa, transb, alpha, htool_b200::DeviceHMatrix<T, U>(A), in, beta, out, mu);
}

// ---- column-major multi-RHS product in cluster numbering: add_hmatrix_matrix_product.hpp:26-77 (transposes around the
// row-major kernel, as the reference does); transb != 'N' and every other output type stay with the reference ------------
template <typename T, typename U>
void b200_internal_add_hmatrix_matrix_product(char transa, char transb, T alpha, const HMatrix<T, U> &A, const Matrix<T> &B, T beta, Matrix<T> &C) {
    if (transb != 'N') {
        internal_add_hmatrix_matrix_product(transa, transb, alpha, A, B, beta, C);
        return;
    }
    Matrix<T> Bt(B.nb_cols(), B.nb_rows()), Ct(C.nb_cols(), C.nb_rows());
    transpose(B, Bt);
    transpose(C, Ct);
    htool_b200::internal_add_hmatrix_matrix_product_row_major(transa, 'N', alpha, htool_b200::DeviceHMatrix<T, U>(A), Bt.data(), beta, Ct.data(), C.nb_cols());
    transpose(Ct, C);
}
template <typename... Args>
void b200_internal_add_hmatrix_matrix_product(Args &&...args) { internal_add_hmatrix_matrix_product(std::forward<Args>(args)...); }

} // namespace htool

#define internal_add_hmatrix_vector_product b200_internal_add_hmatrix_vector_product
#define openmp_internal_add_hmatrix_vector_product b200_openmp_internal_add_hmatrix_vector_product
#define openmp_internal_add_hmatrix_matrix_product_row_major b200_openmp_internal_add_hmatrix_matrix_product_row_major
#define sequential_internal_add_hmatrix_matrix_product_row_major b200_sequential_internal_add_hmatrix_matrix_product_row_major
#define internal_add_hmatrix_matrix_product b200_internal_add_hmatrix_matrix_product

#ifdef REFTEST_COMPLEX
#    include "tests/functional_tests/hmatrix/hmatrix_product/test_hmatrix_product_complex_double.cpp"
#else
#    include "tests/functional_tests/hmatrix/hmatrix_product/test_hmatrix_product_double.cpp"
#endif
