/* oracle/hmat_oracle_impl.h — included twice by hmat_oracle.c with T = double and T = double complex.
 * TEST INFRASTRUCTURE (see hmat_oracle.h). Every function cites the reference lines it restates. */

/* y <- alpha*op(A)*x + beta*y, A m x n column-major, lda = m: what Blas::gemv computes for
 * add_matrix_vector_product (matrix/linalg/add_matrix_vector_product.hpp:10-23). */
static void FN(gemv)(char trans, int m, int n, T alpha, const T *A, const T *x, T beta, T *y) {
    if (trans == 'N') {
        for (int i = 0; i < m; i++) {
            T acc = 0;
            for (int j = 0; j < n; j++)
                acc += A[i + (size_t)m * j] * x[j];
            y[i] = alpha * acc + (beta == (T)0 ? (T)0 : beta * y[i]);
        }
    } else {
        for (int j = 0; j < n; j++) {
            T acc = 0;
            for (int i = 0; i < m; i++)
                acc += (trans == 'C' ? CONJ(A[i + (size_t)m * j]) : A[i + (size_t)m * j]) * x[i];
            y[j] = alpha * acc + (beta == (T)0 ? (T)0 : beta * y[j]);
        }
    }
}

/* symv / hemv on the UPLO triangle of a fully stored n x n block
 * (add_symmetric_matrix_vector_product :26-35, add_hermitian_matrix_vector_product :43-52).
 * hermitian != 0: mirrored entries are conjugated and the imaginary part of the diagonal is ignored,
 * as BLAS zhemv does. */
static void FN(symv)(char uplo, int hermitian, int n, T alpha, const T *A, const T *x, T beta, T *y) {
    for (int i = 0; i < n; i++) {
        T acc = 0;
        for (int j = 0; j < n; j++) {
            T a;
            int stored_ij = (uplo == 'L') ? (i >= j) : (i <= j);
            if (i == j)
                a = hermitian ? (T)REAL(A[i + (size_t)n * i]) : A[i + (size_t)n * i];
            else if (stored_ij)
                a = A[i + (size_t)n * j];
            else
                a = hermitian ? CONJ(A[j + (size_t)n * i]) : A[j + (size_t)n * i];
            acc += a * x[j];
        }
        y[i] = alpha * acc + (beta == (T)0 ? (T)0 : beta * y[i]);
    }
}

/* add_lrmat_vector_product, hmatrix/lrmat/linalg/add_lrmat_vector_product.hpp:9-24 */
static void FN(lrmat_vec)(char trans, int m, int n, int r, T alpha, const T *U, const T *V, const T *in, T beta, T *out) {
    if (r == 0)
        return; /* :11 — product skipped, out untouched */
    T *a = (T *)malloc(sizeof(T) * (size_t)r);
    if (trans == 'N') {
        FN(gemv)('N', r, n, (T)1, V, in, (T)0, a);  /* a = V in  */
        FN(gemv)('N', m, r, alpha, U, a, beta, out); /* out = alpha U a + beta out */
    } else {
        FN(gemv)(trans, m, r, (T)1, U, in, (T)0, a);
        FN(gemv)(trans, r, n, alpha, V, a, beta, out);
    }
    free(a);
}

/* internal_add_hmatrix_vector_product (leaf dispatch), add_hmatrix_vector_product.hpp:17-54 */
static void FN(leaf_vec)(const htb_leaf *l, char trans, T alpha, const T *in, T beta, T *out) {
    if (l->rank < 0) {
        const T *A = (const T *)l->data0;
        if (l->flags & HTB_LEAF_DIAG_SYMMETRIC)
            FN(symv)((l->flags & HTB_LEAF_UPLO_UPPER) ? 'U' : 'L', 0, l->nb_rows, alpha, A, in, beta, out);
        else if (l->flags & HTB_LEAF_DIAG_HERMITIAN)
            FN(symv)((l->flags & HTB_LEAF_UPLO_UPPER) ? 'U' : 'L', IS_COMPLEX, l->nb_rows, alpha, A, in, beta, out);
        else
            FN(gemv)(trans, l->nb_rows, l->nb_cols, alpha, A, in, beta, out);
    } else {
        FN(lrmat_vec)(trans, l->nb_rows, l->nb_cols, l->rank, alpha, (const T *)l->data0, (const T *)l->data1, in, beta, out);
    }
}

/* sequential_internal_add_hmatrix_vector_product, add_hmatrix_vector_product.hpp:57-104 */
static void FN(hmat_vec)(const htb_hmatrix_desc *d, char trans, T alpha, const T *in, T beta, T *out) {
    int out_size   = trans == 'N' ? d->nb_rows : d->nb_cols; /* :64, :75 */
    char trans_sym = d->symmetry_for_leaves == 'S' ? 'T' : 'C'; /* :69 */
    if (trans != 'N')
        trans_sym = 'N'; /* :80 */
    if (beta != (T)1)    /* :83-86 */
        for (int i = 0; i < out_size; i++)
            out[i] = (beta == (T)0) ? (T)0 : beta * out[i];
    for (int64_t b = 0; b < d->nb_leaves; b++) { /* :90-94 */
        const htb_leaf *l = &d->leaves[b];
        int input_offset  = trans == 'N' ? l->col_offset : l->row_offset;
        int output_offset = trans == 'N' ? l->row_offset : l->col_offset;
        FN(leaf_vec)(l, trans, alpha, in + input_offset, (T)1, out + output_offset);
    }
    /* local_output_offset - local_input_offset of the root block (:67-68, :78-79) */
    int shift = trans == 'N' ? d->row_offset - d->col_offset : d->col_offset - d->row_offset;
    if (d->symmetry_for_leaves != 'N') { /* :97-103: leaves_for_symmetry, in/out offsets swapped */
        for (int64_t b = 0; b < d->nb_leaves; b++) {
            const htb_leaf *l = &d->leaves[b];
            if (!(l->flags & HTB_LEAF_APPLY_TRANSPOSED_TOO))
                continue;
            int input_offset  = trans == 'N' ? l->col_offset : l->row_offset;
            int output_offset = trans == 'N' ? l->row_offset : l->col_offset;
            /* leaf offsets are relative to the root block; the reference indexes with absolute offsets
             * (in + output_offset - local_input_offset, out + input_offset - local_output_offset), which
             * differ by the root's own target/source offsets for a row strip */
            FN(leaf_vec)(l, trans_sym, alpha, in + output_offset + shift, (T)1, out + input_offset - shift);
        }
    }
}

/* C(no x mu) <- alpha*op(A)*B(ni x mu) + beta*C with ROW-MAJOR B, C: what the Ct = Bt At gemm of
 * add_matrix_matrix_product_row_major computes (matrix/linalg/add_matrix_matrix_product_row_major.hpp:23-84) */
static void FN(gemm_rm)(char trans, int m, int n, T alpha, const T *A, const T *B, T beta, T *C, int mu) {
    int no = trans == 'N' ? m : n, ni = trans == 'N' ? n : m;
    for (int i = 0; i < no; i++)
        for (int c = 0; c < mu; c++) {
            T acc = 0;
            for (int j = 0; j < ni; j++) {
                T a = trans == 'N' ? A[i + (size_t)m * j] : (trans == 'C' ? CONJ(A[j + (size_t)m * i]) : A[j + (size_t)m * i]);
                acc += a * B[(size_t)j * mu + c];
            }
            C[(size_t)i * mu + c] = alpha * acc + (beta == (T)0 ? (T)0 : beta * C[(size_t)i * mu + c]);
        }
}

/* add_symmetric_/add_hermitian_matrix_matrix_product_row_major with side 'L' (:87-139): C = alpha*A*B + beta*C
 * where A is the symmetric / Hermitian matrix given by its UPLO triangle */
static void FN(symm_rm)(char uplo, int hermitian, int n, T alpha, const T *A, const T *B, T beta, T *C, int mu) {
    for (int i = 0; i < n; i++)
        for (int c = 0; c < mu; c++) {
            T acc = 0;
            for (int j = 0; j < n; j++) {
                T a;
                int stored_ij = (uplo == 'L') ? (i >= j) : (i <= j);
                if (i == j)
                    a = hermitian ? (T)REAL(A[i + (size_t)n * i]) : A[i + (size_t)n * i];
                else if (stored_ij)
                    a = A[i + (size_t)n * j];
                else
                    a = hermitian ? CONJ(A[j + (size_t)n * i]) : A[j + (size_t)n * i];
                acc += a * B[(size_t)j * mu + c];
            }
            C[(size_t)i * mu + c] = alpha * acc + (beta == (T)0 ? (T)0 : beta * C[(size_t)i * mu + c]);
        }
}

/* add_lrmat_matrix_product_row_major, hmatrix/lrmat/linalg/add_lrmat_matrix_product_row_major.hpp:11-27 */
static void FN(lrmat_mat)(char trans, int m, int n, int r, T alpha, const T *U, const T *V, const T *in, T beta, T *out, int mu) {
    if (r == 0)
        return;
    T *a = (T *)malloc(sizeof(T) * (size_t)r * mu);
    if (trans == 'N') {
        FN(gemm_rm)('N', r, n, (T)1, V, in, (T)0, a, mu);
        FN(gemm_rm)('N', m, r, alpha, U, a, beta, out, mu);
    } else {
        FN(gemm_rm)(trans, m, r, (T)1, U, in, (T)0, a, mu);
        FN(gemm_rm)(trans, r, n, alpha, V, a, beta, out, mu);
    }
    free(a);
}

/* internal_add_hmatrix_matrix_product_row_major (leaf dispatch), add_hmatrix_matrix_product_row_major.hpp:18-55 */
static void FN(leaf_mat)(const htb_leaf *l, char trans, T alpha, const T *in, T beta, T *out, int mu) {
    if (l->rank < 0) {
        const T *A = (const T *)l->data0;
        if (l->flags & HTB_LEAF_DIAG_SYMMETRIC)
            FN(symm_rm)((l->flags & HTB_LEAF_UPLO_UPPER) ? 'U' : 'L', 0, l->nb_rows, alpha, A, in, beta, out, mu);
        else if (l->flags & HTB_LEAF_DIAG_HERMITIAN)
            FN(symm_rm)((l->flags & HTB_LEAF_UPLO_UPPER) ? 'U' : 'L', IS_COMPLEX, l->nb_rows, alpha, A, in, beta, out, mu);
        else
            FN(gemm_rm)(trans, l->nb_rows, l->nb_cols, alpha, A, in, beta, out, mu);
    } else {
        FN(lrmat_mat)(trans, l->nb_rows, l->nb_cols, l->rank, alpha, (const T *)l->data0, (const T *)l->data1, in, beta, out, mu);
    }
}

/* sequential_internal_add_hmatrix_matrix_product_row_major, add_hmatrix_matrix_product_row_major.hpp:58-109:
 * beta scaling of C, leaf loop into a zeroed temp with alpha = 1, leaves_for_symmetry with swapped
 * offsets, then C += alpha*temp (:108) */
static void FN(hmat_mat)(const htb_hmatrix_desc *d, char trans, T alpha, const T *B, T beta, T *C, int mu) {
    size_t out_size = (size_t)(trans == 'N' ? d->nb_rows : d->nb_cols) * mu;
    char trans_sym  = d->symmetry_for_leaves == 'S' ? 'T' : 'C';
    if (trans != 'N')
        trans_sym = 'N';
    if (beta != (T)1)
        for (size_t i = 0; i < out_size; i++)
            C[i] = (beta == (T)0) ? (T)0 : beta * C[i];
    T *temp = (T *)calloc(out_size ? out_size : 1, sizeof(T));
    for (int64_t b = 0; b < d->nb_leaves; b++) {
        const htb_leaf *l = &d->leaves[b];
        int input_offset  = trans == 'N' ? l->col_offset : l->row_offset;
        int output_offset = trans == 'N' ? l->row_offset : l->col_offset;
        FN(leaf_mat)(l, trans, (T)1, B + (size_t)input_offset * mu, (T)1, temp + (size_t)output_offset * mu, mu);
    }
    int shift = trans == 'N' ? d->row_offset - d->col_offset : d->col_offset - d->row_offset;
    if (d->symmetry_for_leaves != 'N') {
        for (int64_t b = 0; b < d->nb_leaves; b++) {
            const htb_leaf *l = &d->leaves[b];
            if (!(l->flags & HTB_LEAF_APPLY_TRANSPOSED_TOO))
                continue;
            int input_offset  = trans == 'N' ? l->col_offset : l->row_offset;
            int output_offset = trans == 'N' ? l->row_offset : l->col_offset;
            FN(leaf_mat)(l, trans_sym, (T)1, B + (size_t)(output_offset + shift) * mu, (T)1, temp + (size_t)(input_offset - shift) * mu, mu);
        }
    }
    for (size_t i = 0; i < out_size; i++)
        C[i] += alpha * temp[i];
    free(temp);
}
