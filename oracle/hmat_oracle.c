/* oracle/hmat_oracle.c — TEST INFRASTRUCTURE (see hmat_oracle.h): the reference's sequential product
 * restated in plain C over flattened leaves, double and complex<double>. */
#include "hmat_oracle.h"
#include <complex.h>
#include <stdlib.h>
#include <string.h>

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

#define T double
#define FN(name) CAT(d_, name)
#define CONJ(x) (x)
#define REAL(x) (x)
#define IS_COMPLEX 0
#include "hmat_oracle_impl.h"
#undef T
#undef FN
#undef CONJ
#undef REAL
#undef IS_COMPLEX

#define T double complex
#define FN(name) CAT(z_, name)
#define CONJ(x) conj(x)
#define REAL(x) creal(x)
#define IS_COMPLEX 1
#include "hmat_oracle_impl.h"

static int rejected(const htb_hmatrix_desc *d, char trans) {
    /* add_hmatrix_vector_product.hpp:59-62 */
    return (trans == 'T' && d->symmetry_for_leaves == 'H') || (trans == 'C' && d->symmetry_for_leaves == 'S');
}

int oracle_add_vector_product(const htb_hmatrix_desc *d, char trans, const void *alpha, const void *in, const void *beta, void *out) {
    if (rejected(d, trans))
        return 2;
    if (d->dtype == HTB_DOUBLE)
        d_hmat_vec(d, trans, *(const double *)alpha, (const double *)in, *(const double *)beta, (double *)out);
    else
        z_hmat_vec(d, trans, *(const double complex *)alpha, (const double complex *)in, *(const double complex *)beta, (double complex *)out);
    return 0;
}

int oracle_add_matrix_product_row_major(const htb_hmatrix_desc *d, char trans, const void *alpha, const void *in, const void *beta, void *out, int mu) {
    if (rejected(d, trans))
        return 2;
    if (d->dtype == HTB_DOUBLE)
        d_hmat_mat(d, trans, *(const double *)alpha, (const double *)in, *(const double *)beta, (double *)out, mu);
    else
        z_hmat_mat(d, trans, *(const double complex *)alpha, (const double complex *)in, *(const double complex *)beta, (double complex *)out, mu);
    return 0;
}
