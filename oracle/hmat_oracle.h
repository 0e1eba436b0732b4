/* oracle/hmat_oracle.h — TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's H-matrix product on FLATTENED leaves (the htb_leaf list of
 * include/htool_b200.h). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call
 * it, and only as the checker. Parity is PINNED: tests/test_oracle.py checks it against the golden
 * vectors in tests/golden/ (produced by the unmodified reference through oracle/_ref, script
 * tools/make_golden.py) and, when oracle/_ref is present, against the reference run live.
 */
#ifndef HTB_HMAT_ORACLE_H
#define HTB_HMAT_ORACLE_H
#include "../include/htool_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* follows sequential_internal_add_hmatrix_vector_product, add_hmatrix_vector_product.hpp:57-104.
 * returns 0, or 2 for the combinations the reference rejects (:59-62). */
int oracle_add_vector_product(const htb_hmatrix_desc *desc, char trans, const void *alpha, const void *in, const void *beta, void *out);
/* follows sequential_internal_add_hmatrix_matrix_product_row_major, add_hmatrix_matrix_product_row_major.hpp:58-109 */
int oracle_add_matrix_product_row_major(const htb_hmatrix_desc *desc, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu);
#ifdef __cplusplus
}
#endif
#endif
