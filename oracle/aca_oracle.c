/* oracle/aca_oracle.c — TEST INFRASTRUCTURE (never linked, loaded or executed by the product).
 *
 * Plain-C restatement of the reference's default compressor for one admissible block of a REAL kernel function:
 *   sympartialACA::copy_low_rank_approximation   include/htool/hmatrix/lrmat/sympartialACA.hpp:41-216
 * with the reference's stopping criterion (reqrank < 0), on the analytic generators of
 *   include/htool/testing/generator_test.hpp:155-161 (kernel 0: 1 / (4 pi r)) and :180-187 (kernel 1: 1 / (1e-5 + 4 pi r))
 * evaluated at points given in CLUSTER numbering (what InternalGenerator::copy_submatrix resolves through the permutation).
 *
 * Pinned by tests/test_aca_oracle.py against the reference run live (oracle/_ref: the low-rank leaves HMatrixTreeBuilder
 * produces with its default sympartialACA: same ranks, same U and V) and against the fixtures of tests/golden/aca_*.npz
 * generated from the unmodified reference by tools/make_golden_aca.py.
 *
 * The only freedom the reference leaves is its BLAS: daxpy computes y += a x either with a rounded product and a rounded
 * sum or with one fused multiply-add, depending on the library build and the CPU it dispatches on. `fma_axpy` selects the
 * variant; everything else (the dot products of wrapper_blas.hpp:152-157, the pivot scans, the generator) is plain C++ in the
 * reference and plain C here, compiled without contraction (-ffp-contract=off). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static double kernel_value(int kernel, const double *a, const double *b) {
    /* generator_test.hpp:155-161 / :180-187 (ref_harness.hpp KernelGenerator): (target point, source point) */
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    const double r  = sqrt(dx * dx + dy * dy + dz * dz);
    const double fpr = (4 * M_PI) * r;
    return kernel == 0 ? 1. / fpr : 1. / (1e-5 + fpr);
}

static void axpy(int n, double coef, const double *x, double *y, int fma_axpy) {
    if (fma_axpy)
        for (int i = 0; i < n; i++)
            y[i] = fma(coef, x[i], y[i]);
    else
        for (int i = 0; i < n; i++)
            y[i] += coef * x[i];
}

static double dot(int n, const double *x, const double *y) { /* wrapper_blas.hpp:152-157 */
    double sum = 0.;
    for (int i = 0; i < n; i++)
        sum += x[i] * y[i];
    return sum;
}

/* One block: rows [lrow, lrow + M) x columns [lcol, lcol + N) of the root block; row_offset / col_offset are the GLOBAL
 * offsets of its clusters (they decide which dimension is "1", sympartialACA.hpp:46-66). target_points / source_points:
 * 3 doubles per row / column of the ROOT block. U (M x rank, column-major) and V (rank x N, column-major) must hold
 * max_rank terms. pivots (optional, 2 * max_rank ints): I1 then I2 of every accepted term.
 * Returns the rank (> 0), or -1 when the reference reports a failure (the leaf becomes dense), or -2 when max_rank is too small. */
int oracle_sympartial_aca(int kernel, const double *target_points, const double *source_points, int M, int N, int row_offset, int col_offset, int lrow, int lcol, double epsilon, int fma_axpy, int max_rank, double *U, double *V,
                          int *pivots) {
    const int direct = row_offset >= col_offset; /* :46 */
    const int n1 = direct ? M : N, n2 = direct ? N : M;
    const double *p1 = direct ? target_points + 3 * (size_t)lrow : source_points + 3 * (size_t)lcol;
    const double *p2 = direct ? source_points + 3 * (size_t)lcol : target_points + 3 * (size_t)lrow;
    int I1 = 0, I2 = 0, q = 0;
    double **uu = calloc((size_t)max_rank + 1, sizeof(double *)), **vv = calloc((size_t)max_rank + 1, sizeof(double *));
    char *visited_1 = calloc((size_t)n1 + 1, 1), *visited_2 = calloc((size_t)n2 + 1, 1);
    double *u1 = malloc(sizeof(double) * ((size_t)n2 + 1)), *u2 = malloc(sizeof(double) * ((size_t)n1 + 1));
    int nterms = 0, status = 0;
    double frob = 0., aux = 0.;

    while (q == 0 || sqrt(aux / frob) > epsilon) { /* :97 (reqrank < 0) */
        q += 1;
        if ((int64_t)q * ((int64_t)n1 + n2) > (int64_t)n1 * n2) { /* :102 */
            q = -1;
            break;
        }
        if (nterms == max_rank) {
            status = -2;
            break;
        }
        /* row I1 of the block (:104-110); the generator is called as (target, source) in both orientations */
        for (int k = 0; k < n2; k++)
            u1[k] = direct ? kernel_value(kernel, p1 + 3 * (size_t)I1, p2 + 3 * (size_t)k) : kernel_value(kernel, p2 + 3 * (size_t)k, p1 + 3 * (size_t)I1);
        for (int j = 0; j < nterms; j++)
            axpy(n2, -uu[j][I1], vv[j], u1, fma_axpy);
        double pivot = 0., tmp; /* :112-122: the LAST largest unvisited entry */
        for (int k = 0; k < n2; k++) {
            if (visited_2[k])
                continue;
            tmp = fabs(u1[k]);
            if (tmp < pivot)
                continue;
            pivot = tmp;
            I2    = k;
        }
        visited_1[I1]      = 1;
        const double gamma = 1. / u1[I2]; /* :124 */
        if (fabs(u1[I2]) > 1e-15) {       /* :128 */
            for (int k = 0; k < n1; k++)
                u2[k] = direct ? kernel_value(kernel, p1 + 3 * (size_t)k, p2 + 3 * (size_t)I2) : kernel_value(kernel, p2 + 3 * (size_t)I2, p1 + 3 * (size_t)k);
            for (int k = 0; k < nterms; k++)
                axpy(n1, -vv[k][I2], uu[k], u2, fma_axpy);
            for (int k = 0; k < n1; k++) /* :142, u2 *= gamma (misc/misc.hpp operator*=: element by element) */
                u2[k] *= gamma;
            if (pivots) {
                pivots[2 * nterms]     = I1;
                pivots[2 * nterms + 1] = I2;
            }
            pivot = 0.;
            for (int k = 0; k < n1; k++) { /* :143-152 */
                if (visited_1[k])
                    continue;
                tmp = fabs(u2[k]);
                if (tmp < pivot)
                    continue;
                pivot = tmp;
                I1    = k;
            }
            visited_2[I2] = 1;
            /* error estimator (:156-168) */
            double frob_aux = 0.;
            aux             = fabs(dot(n1, u2, u2)) * fabs(dot(n2, u1, u1));
            for (int j = 0; j < nterms; j++)
                frob_aux += dot(n2, u1, vv[j]) * dot(n1, u2, uu[j]);
            frob += aux + 2 * frob_aux;
            uu[nterms] = malloc(sizeof(double) * (size_t)n1);
            vv[nterms] = malloc(sizeof(double) * (size_t)n2);
            memcpy(uu[nterms], u2, sizeof(double) * (size_t)n1);
            memcpy(vv[nterms], u1, sizeof(double) * (size_t)n2);
            nterms++;
        } else { /* :184-193: a zero row */
            q -= 1;
            if (q == 0)
                q = -1;
            break;
        }
    }
    if (status == 0 && q > 0) { /* :198-222: U = [uu_k] and V = [vv_k^T], or the other way round */
        for (int k = 0; k < q; k++) {
            const double *cu = direct ? uu[k] : vv[k], *cv = direct ? vv[k] : uu[k];
            memcpy(U + (size_t)k * M, cu, sizeof(double) * (size_t)M);
            for (int j = 0; j < N; j++)
                V[k + (size_t)j * q] = cv[j];
        }
    }
    for (int k = 0; k < nterms; k++) {
        free(uu[k]);
        free(vv[k]);
    }
    free(uu), free(vv), free(visited_1), free(visited_2), free(u1), free(u2);
    return status ? status : q;
}

/* ---- complex<double> kernel functions -------------------------------------------------------------------------------
 * Same algorithm on std::complex<double> (the reference's template instantiated for complex coefficients): |.| is
 * std::abs = cabs, gamma = 1 / pivot and u2 *= gamma are the compiler's complex division / product (libgcc __divdc3 /
 * inline product, the same routines libstdc++ ends up in), the dot products conjugate their first argument
 * (wrapper_blas.hpp:152-157, conj_if_complex), frob takes the real part of the cross terms (:166).
 * NOT bit-pinned: zaxpy belongs to the BLAS (its product / sum order and FMA use are the library's), and the Helmholtz
 * kernel goes through the C library's sincos. Pinned to rounding: same pivots and ranks, factors to ~1e-13
 * (tests/test_aca_oracle.py). Kernel ids = oracle/ref/ref_harness.hpp: 2 (1+i)/(1e-5+4 pi r), 3 Hermitian, 4 Helmholtz, 5 (1+i)/(4 pi r). */
#include <complex.h>

static double complex kernel_value_z(int kernel, const double *a, const double *b, double k) {
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    const double r   = sqrt(dx * dx + dy * dy + dz * dz);
    const double fpr = (4 * M_PI) * r;
    if (kernel == 5) {
        const double v = 1. / fpr;
        return v + v * I;
    }
    if (kernel == 3) {
        const double d = 1e-5 + fpr, s = dx > 0 ? 1. : (dx < 0 ? -1. : 0.);
        return 1. / d + (s / d) * I;
    }
    if (kernel == 4) {
        if (r < 1e-12)
            return 1. / ((4 * M_PI) * 1e-3) + (k / (4 * M_PI)) * I;
        const double complex e = cexp(0. + (k * r) * I);
        return creal(e) / fpr + (cimag(e) / fpr) * I;
    }
    const double v = 1. / (1e-5 + fpr);
    return v + v * I;
}

static void zaxpy(int n, double complex coef, const double complex *x, double complex *y, int fma_axpy) {
    const double ar = creal(coef), ai = cimag(coef);
    for (int i = 0; i < n; i++) {
        const double xr = creal(x[i]), xi = cimag(x[i]);
        double yr = creal(y[i]), yi = cimag(y[i]);
        if (fma_axpy) { /* the order of OpenBLAS' FMA kernels: y += ar * x, then y += ai * (i x) */
            yr = fma(ar, xr, yr);
            yi = fma(ar, xi, yi);
            yr = fma(-ai, xi, yr);
            yi = fma(ai, xr, yi);
        } else { /* OpenBLAS' SSE2 kernel: y += ar * x, then y += ai * (i x), every product and sum rounded */
            yr = (yr + ar * xr) - ai * xi;
            yi = (yi + ar * xi) + ai * xr;
        }
        y[i] = yr + yi * I;
    }
}

static double complex zdot(int n, const double complex *x, const double complex *y) {
    double complex sum = 0.;
    for (int i = 0; i < n; i++)
        sum += conj(x[i]) * y[i];
    return sum;
}

/* As oracle_sympartial_aca; U, V and the points as there, coefficients complex (re, im interleaved). */
int oracle_sympartial_aca_z(int kernel, double wavenumber, const double *target_points, const double *source_points, int M, int N, int row_offset, int col_offset, int lrow, int lcol, double epsilon, int fma_axpy, int max_rank,
                            double complex *U, double complex *V, int *pivots) {
    const int direct = row_offset >= col_offset;
    const int n1 = direct ? M : N, n2 = direct ? N : M;
    const double *p1 = direct ? target_points + 3 * (size_t)lrow : source_points + 3 * (size_t)lcol;
    const double *p2 = direct ? source_points + 3 * (size_t)lcol : target_points + 3 * (size_t)lrow;
    int I1 = 0, I2 = 0, q = 0;
    double complex **uu = calloc((size_t)max_rank + 1, sizeof(double complex *)), **vv = calloc((size_t)max_rank + 1, sizeof(double complex *));
    char *visited_1 = calloc((size_t)n1 + 1, 1), *visited_2 = calloc((size_t)n2 + 1, 1);
    double complex *u1 = malloc(sizeof(double complex) * ((size_t)n2 + 1)), *u2 = malloc(sizeof(double complex) * ((size_t)n1 + 1));
    int nterms = 0, status = 0;
    double frob = 0., aux = 0.;

    while (q == 0 || sqrt(aux / frob) > epsilon) {
        q += 1;
        if ((int64_t)q * ((int64_t)n1 + n2) > (int64_t)n1 * n2) {
            q = -1;
            break;
        }
        if (nterms == max_rank) {
            status = -2;
            break;
        }
        for (int k = 0; k < n2; k++)
            u1[k] = direct ? kernel_value_z(kernel, p1 + 3 * (size_t)I1, p2 + 3 * (size_t)k, wavenumber) : kernel_value_z(kernel, p2 + 3 * (size_t)k, p1 + 3 * (size_t)I1, wavenumber);
        for (int j = 0; j < nterms; j++)
            zaxpy(n2, -uu[j][I1], vv[j], u1, fma_axpy);
        double pivot = 0., tmp;
        for (int k = 0; k < n2; k++) {
            if (visited_2[k])
                continue;
            tmp = cabs(u1[k]);
            if (tmp < pivot)
                continue;
            pivot = tmp;
            I2    = k;
        }
        visited_1[I1]              = 1;
        const double complex gamma = (1. + 0. * I) / u1[I2];
        if (cabs(u1[I2]) > 1e-15) {
            for (int k = 0; k < n1; k++)
                u2[k] = direct ? kernel_value_z(kernel, p1 + 3 * (size_t)k, p2 + 3 * (size_t)I2, wavenumber) : kernel_value_z(kernel, p2 + 3 * (size_t)I2, p1 + 3 * (size_t)k, wavenumber);
            for (int k = 0; k < nterms; k++)
                zaxpy(n1, -vv[k][I2], uu[k], u2, fma_axpy);
            for (int k = 0; k < n1; k++)
                u2[k] = u2[k] * gamma;
            if (pivots) {
                pivots[2 * nterms]     = I1;
                pivots[2 * nterms + 1] = I2;
            }
            pivot = 0.;
            for (int k = 0; k < n1; k++) {
                if (visited_1[k])
                    continue;
                tmp = cabs(u2[k]);
                if (tmp < pivot)
                    continue;
                pivot = tmp;
                I1    = k;
            }
            visited_2[I2] = 1;
            double complex frob_aux = 0.;
            aux                     = cabs(zdot(n1, u2, u2)) * cabs(zdot(n2, u1, u1));
            for (int j = 0; j < nterms; j++)
                frob_aux += zdot(n2, u1, vv[j]) * zdot(n1, u2, uu[j]);
            frob += aux + 2 * creal(frob_aux);
            uu[nterms] = malloc(sizeof(double complex) * (size_t)n1);
            vv[nterms] = malloc(sizeof(double complex) * (size_t)n2);
            memcpy(uu[nterms], u2, sizeof(double complex) * (size_t)n1);
            memcpy(vv[nterms], u1, sizeof(double complex) * (size_t)n2);
            nterms++;
        } else {
            q -= 1;
            if (q == 0)
                q = -1;
            break;
        }
    }
    if (status == 0 && q > 0) {
        for (int k = 0; k < q; k++) {
            const double complex *cu = direct ? uu[k] : vv[k], *cv = direct ? vv[k] : uu[k];
            memcpy(U + (size_t)k * M, cu, sizeof(double complex) * (size_t)M);
            for (int j = 0; j < N; j++)
                V[k + (size_t)j * q] = cv[j];
        }
    }
    for (int k = 0; k < nterms; k++) {
        free(uu[k]);
        free(vv[k]);
    }
    free(uu), free(vv), free(visited_1), free(visited_2), free(u1), free(u2);
    return status ? status : q;
}
