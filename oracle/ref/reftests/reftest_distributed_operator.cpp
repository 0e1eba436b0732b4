// oracle/ref/reftests/reftest_distributed_operator.cpp — TEST INFRASTRUCTURE. The reference's OWN test program
//   tests/functional_tests/distributed_operator/test_distributed_operator_product_{double,complex_double}.cpp
//   (+ test_distributed_operator.hpp:47-384, the 32 checks per configuration: global-to-global / local-to-local, vector /
//   matrix / row-major, with and without sum, user and partition numbering)
// compiled UNMODIFIED where it lies under /root/reference, with the B200 twins substituted for the reference's two
// H-matrix operator adapters at include time:
//   1. the reference's own adapter headers are included first, so their include guards are set and the class templates
//      htool::RestrictedGlobalToLocalHMatrix / htool::LocalToLocalHMatrix keep their names;
//   2. the twins (htool_b200/operators.hpp) are aliased into namespace htool;
//   3. two #defines make every LATER mention of the reference's class names — DefaultApproximationBuilder /
//      DefaultLocalApproximationBuilder in distributed_operator/utility.hpp:37-96 and the off-diagonal operators the test
//      builds itself (test_distributed_operator.hpp:462-463) — mean the twins;
//   4. the reference's test main() follows.
// Every H-matrix product of the test therefore runs on the GPU through the C ABI; the dense-matrix operator variants and
// all the DistributedOperator linalg around the operators stay the reference's code. Exit code 0 = the reference's own
// tolerances hold with the twins. Selected with -DREFTEST_COMPLEX for the complex<double> program.
#include <htool/distributed_operator/implementations/global_to_local_operators/hmatrix.hpp>
#include <htool/distributed_operator/implementations/local_to_local_operators/hmatrix.hpp>

#include <htool_b200/operators.hpp>

namespace htool {
template <typename CoefficientPrecision, typename CoordinatePrecision = underlying_type<CoefficientPrecision>>
using B200RestrictedGlobalToLocalHMatrix = htool_b200::RestrictedGlobalToLocalHMatrix<CoefficientPrecision, CoordinatePrecision>;
template <typename CoefficientPrecision, typename CoordinatePrecision = underlying_type<CoefficientPrecision>>
using B200LocalToLocalHMatrix = htool_b200::LocalToLocalHMatrix<CoefficientPrecision, CoordinatePrecision>;
} // namespace htool

#define RestrictedGlobalToLocalHMatrix B200RestrictedGlobalToLocalHMatrix
#define LocalToLocalHMatrix B200LocalToLocalHMatrix

#ifdef REFTEST_COMPLEX
#    include "tests/functional_tests/distributed_operator/test_distributed_operator_product_complex_double.cpp"
#else
#    include "tests/functional_tests/distributed_operator/test_distributed_operator_product_double.cpp"
#endif
