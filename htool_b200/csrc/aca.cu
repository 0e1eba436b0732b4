// htool_b200/csrc/aca.cu — batched adaptive cross approximation on the device: the admissible blocks of the block cluster tree
// are compressed on the GPU, straight from the kernel function and the points, and their factors go into the leaf store
// without ever existing on the host (second step; the dense leaves are generate.cu).
//
// What it replaces in the reference: sympartialACA::copy_low_rank_approximation
// (include/htool/hmatrix/lrmat/sympartialACA.hpp:41-216), called per admissible block from HMatrix::compute_low_rank_data
// (hmatrix/hmatrix.hpp:228-237) inside HMatrixTreeBuilder::openmp_compute_blocks (hmatrix/tree_builder/tree_builder.hpp:604-626):
// 1.0 M blocks / 2.32 G coefficients / 11 s on 16 host cores at N = 1e6.
//
// The arithmetic is the reference's, operation for operation, so that a real kernel function produces the SAME pivots, the
// same ranks and bit-identical factors:
//   * the generator values are those of kernel_functions.cuh (bit-identical to the host generators);
//   * the residual updates are the reference's axpy chain in term order (u -= uu_j[I1] vv_j for j = 0, 1, ...: :107-110,
//     :138-141), one rounded product + one rounded sum per term (fma_axpy = 1: one fused operation, for FMA BLAS builds);
//   * the pivot searches take the LAST maximum among the unvisited entries (the reference's `if (tmp < pivot) continue`,
//     :114-122, :144-152): a team-wide arg-max with "larger index wins ties";
//   * the dot products of the stopping criterion (:158-166) are the reference's own scalar loop
//     (wrapper_blas.hpp:152-157: sum += x[i] * y[i], in index order): one THREAD per dot product walks its vectors in order —
//     the 2 q dot products of an iteration are independent and run side by side;
//   * complex kernel functions: aca_kernel_z below (same structure, the complex arithmetic of the reference's instantiation);
//   * q (n1 + n2) > n1 n2 is evaluated in 64 bits (the reference's int product overflows from 46341 x 46341 on).
// One CTA ("team") per block; blocks are sorted by size and launched in three classes (512 / 128 / 32 threads). Every new
// term takes a chunk [uu (n1) | vv (n2)] from a bump-allocated pool; the block's chunk offsets are its term table.
#include "aca.cuh"
#include "kernel_functions.cuh"

namespace htb {

namespace {

template <int KERNEL>
__device__ __forceinline__ double entry(bool swapped, const double *p1, const double *p2, int a, int b, double wavenumber) {
    // a: index along dimension 1, b: along dimension 2; the kernel function takes (target point, source point)
    return swapped ? kernel_value<KERNEL>(p2 + 3ll * b, p1 + 3ll * a, wavenumber).re : kernel_value<KERNEL>(p1 + 3ll * a, p2 + 3ll * b, wavenumber).re;
}

template <bool FMA>
__device__ __forceinline__ double axpy1(double coef, double x, double y) {
    return FMA ? fma(coef, x, y) : __dadd_rn(y, __dmul_rn(coef, x));
}

// sum += x[i] * y[i] in index order, every operation rounded (wrapper_blas.hpp:152-157)
__device__ __forceinline__ double dot_in_order(const double *x, const double *y, int len) {
    double s = 0.;
    int i    = 0;
    for (; i + 4 <= len; i += 4) {
        const double x0 = x[i], x1 = x[i + 1], x2 = x[i + 2], x3 = x[i + 3];
        const double y0 = y[i], y1 = y[i + 1], y2 = y[i + 2], y3 = y[i + 3];
        s = __dadd_rn(s, __dmul_rn(x0, y0));
        s = __dadd_rn(s, __dmul_rn(x1, y1));
        s = __dadd_rn(s, __dmul_rn(x2, y2));
        s = __dadd_rn(s, __dmul_rn(x3, y3));
    }
    for (; i < len; i++)
        s = __dadd_rn(s, __dmul_rn(x[i], y[i]));
    return s;
}

// Team-wide arg-max of (value, index): the larger value wins, equal values: the larger index (the reference's scan keeps the
// LAST maximum). index -1 = no candidate. Every thread returns the result.
template <int TS>
__device__ __forceinline__ int team_argmax(double best, int idx, double *s_rv, int *s_ri) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi    = __shfl_xor_sync(0xffffffffu, idx, d);
        if (ob > best || (ob == best && oi > idx)) {
            best = ob;
            idx  = oi;
        }
    }
    if (TS == 32)
        return idx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads(); // (the scratch may still be read by the previous call)
    if (lane == 0) {
        s_rv[warp] = best;
        s_ri[warp] = idx;
    }
    __syncthreads();
    best = s_rv[0], idx = s_ri[0];
#pragma unroll
    for (int w = 1; w < TS / 32; w++) {
        const double ob = s_rv[w];
        const int oi    = s_ri[w];
        if (ob > best || (ob == best && oi > idx)) {
            best = ob;
            idx  = oi;
        }
    }
    return idx;
}

template <int TS>
__device__ __forceinline__ void team_sync() {
    if (TS == 32)
        __syncwarp();
    else
        __syncthreads();
}

// The reference's error estimator for the term (uq, vq) that follows q - 1 accepted terms (sympartialACA.hpp:156-168): 2 q dot
// products in index order, one THREAD each (they are independent), then aux = |<u,u>| |<v,v>| and
// frob += aux + 2 sum_j <v, vv_j> <u, uu_j> accumulated in term order. Every thread returns the same values.
template <int TS>
__device__ __forceinline__ void in_order_estimator(int q, const double *uq, const double *vq, int n1, int n2, const double *pool, const uint32_t *s_off, double *s_dot, int tid, double &aux, double &frob) {
    const int ndots = 2 * q;
    for (int t = tid; t < ndots; t += TS) {
        double d;
        if (t == 0)
            d = dot_in_order(uq, uq, n1);
        else if (t == 1)
            d = dot_in_order(vq, vq, n2);
        else {
            const double *ch = pool + 2ull * s_off[(t - 2) >> 1];
            d                = (t & 1) ? dot_in_order(uq, ch, n1) : dot_in_order(vq, ch + n1, n2);
        }
        s_dot[t] = d;
    }
    team_sync<TS>();
    aux             = __dmul_rn(fabs(s_dot[0]), fabs(s_dot[1]));
    double frob_aux = 0.;
    for (int j = 0; j < q - 1; j++)
        frob_aux = __dadd_rn(frob_aux, __dmul_rn(s_dot[2 + 2 * j], s_dot[3 + 2 * j]));
    frob = __dadd_rn(frob, __dadd_rn(aux, __dmul_rn(2., frob_aux)));
}

template <int TS, int KERNEL, bool FMA>
__global__ void __launch_bounds__(TS, TS == 512 ? 2 : (TS == 128 ? 8 : 28)) aca_kernel(const AcaBlock *__restrict__ blocks, long long first, const double *__restrict__ tp, const double *__restrict__ sp, double wavenumber, double epsilon, AcaPool pool,
                                                 int32_t *__restrict__ rank_out, int dots_mode) {
    constexpr int kCap = aca_max_rank(TS);
    __shared__ uint32_t s_off[kCap];
    __shared__ int s_piv1[kCap], s_piv2[kCap];
    __shared__ double s_dot[2 * kCap + 2];
    __shared__ double s_rv[TS / 32];
    __shared__ int s_ri[TS / 32];
    __shared__ unsigned long long s_chunk;

    const long long bi = first + blockIdx.x;
    const AcaBlock blk = blocks[bi];
    const int tid      = threadIdx.x;
    const bool swapped = blk.swapped != 0;
    const int n1 = swapped ? blk.n : blk.m, n2 = swapped ? blk.m : blk.n;
    const double *p1 = swapped ? sp + 3ll * blk.lcol : tp + 3ll * blk.lrow; // points of dimension 1
    const double *p2 = swapped ? tp + 3ll * blk.lrow : sp + 3ll * blk.lcol;

    // (sympartialACA.hpp:69-92) every thread carries the same copy of the scalar state
    int q = 0, I1 = 0, I2 = 0, nv1 = 0, nv2 = 0;
    double frob = 0., aux = 0.;
    double frob_fast = 0.;   // running Frobenius estimate of the guarded fast path (below)
    bool go_on       = true; // the reference's loop condition (:97), decided at the end of the iteration
    while (go_on) {
        q += 1;
        if (static_cast<long long>(q) * (static_cast<long long>(n1) + n2) > static_cast<long long>(n1) * n2) { // :102: the next rank would not be advantageous
            q = kAcaFailed;
            break;
        }
        if (q > static_cast<int>(blk.term_cap)) {
            q = kAcaRankCap;
            break;
        }
        // a chunk [uu_q | vv_q] for the new term
        const unsigned long long len = (static_cast<unsigned long long>(n1) + n2 + 1ull) & ~1ull;
        team_sync<TS>();
        if (tid == 0) {
            const unsigned long long off = atomicAdd(pool.cursor, len);
            s_chunk                      = off + len <= pool.capacity ? off : ~0ull;
            if (off + len <= pool.capacity)
                s_off[q - 1] = static_cast<uint32_t>(off >> 1);
        }
        team_sync<TS>();
        const unsigned long long off = s_chunk;
        if (off == ~0ull) {
            q = kAcaPoolOverflow;
            break;
        }
        double *uq = pool.pool + off, *vq = uq + n1;

        // row I1 of the residual (:105-110), pivot among the unvisited columns (:112-122)
        double best = 0.;
        int idx     = -1;
        for (int b = tid; b < n2; b += TS) {
            double v = entry<KERNEL>(swapped, p1, p2, I1, b, wavenumber);
            for (int j = 0; j < q - 1; j++) {
                const double *ch = pool.pool + 2ull * s_off[j];
                v                = axpy1<FMA>(-ch[I1], ch[n1 + b], v);
            }
            vq[b]     = v;
            bool seen = false;
            for (int t = 0; t < nv2; t++)
                seen |= s_piv2[t] == b;
            if (!seen) {
                const double tmp = fabs(v);
                if (tmp >= best) {
                    best = tmp;
                    idx  = b;
                }
            }
        }
        idx = team_argmax<TS>(best, idx, s_rv, s_ri);
        if (idx >= 0)
            I2 = idx;
        if (tid == 0)
            s_piv1[nv1] = I1; // :123
        nv1++;
        team_sync<TS>(); // vq written by the team is visible to every thread
        const double pivot = vq[I2];
        if (!(fabs(pivot) > 1e-15)) { // :128 / :185-193: a zero row
            q -= 1;
            if (q == 0)
                q = kAcaFailed;
            break;
        }
        const double gamma = __ddiv_rn(1., pivot); // :124

        // column I2 of the residual, scaled (:130-142), pivot among the unvisited rows (:143-152)
        best = 0.;
        idx  = -1;
        for (int a = tid; a < n1; a += TS) {
            double v = entry<KERNEL>(swapped, p1, p2, a, I2, wavenumber);
            for (int j = 0; j < q - 1; j++) {
                const double *ch = pool.pool + 2ull * s_off[j];
                v                = axpy1<FMA>(-ch[n1 + I2], ch[a], v);
            }
            v         = __dmul_rn(v, gamma);
            uq[a]     = v;
            bool seen = false;
            for (int t = 0; t < nv1; t++)
                seen |= s_piv1[t] == a;
            if (!seen) {
                const double tmp = fabs(v);
                if (tmp >= best) {
                    best = tmp;
                    idx  = a;
                }
            }
        }
        idx = team_argmax<TS>(best, idx, s_rv, s_ri);
        if (idx >= 0)
            I1 = idx;
        if (tid == 0)
            s_piv2[nv2] = I2; // :153
        nv2++;
        team_sync<TS>();

        // error estimator (:156-168)
        if (TS == 32 || dots_mode == 1) {
            // the reference's arithmetic: 2 q dot products, one thread each, every one in index order
            in_order_estimator<TS>(q, uq, vq, n1, n2, pool.pool, s_off, s_dot, tid, aux, frob);
            go_on = __dsqrt_rn(__ddiv_rn(aux, frob)) > epsilon;
        } else {
            // Large blocks: an in-order dot product is ONE serial chain of n additions and the rest of the team waits for
            // it at the barrier (ncu: 5 % issue-active, 134 barrier stalls per issue). The dot products only feed the
            // STOPPING DECISION sqrt(aux / frob) > epsilon — the factors do not depend on them — so the decision is taken
            // from dot products computed by whole warps (lane-strided FMA partial sums + shuffle reduction): they differ
            // from the in-order sums by <= (n / 32 + 5) u |x|.|y|, i.e. the ratio by well under 1e-8 relative for
            // any block the kernel accepts (q <= 128, n <= 2^31). Whenever the fast ratio lies within kGuard (1e-6,
            // relative) of epsilon, or is not finite, the decision is taken instead from the reference's own arithmetic,
            // replayed in order from the first term (rare: the ratio falls by ~10x per iteration). Same ranks as the
            // reference, provably; dots_mode 2 (tests) forces the replay at every iteration.
            constexpr int W = TS / 32;
            const int w = tid >> 5, ln = tid & 31;
            const int ndots = 2 * q;
            for (int t = w; t < ndots; t += W) {
                const double *x, *y;
                int len;
                if (t == 0)
                    x = uq, y = uq, len = n1;
                else if (t == 1)
                    x = vq, y = vq, len = n2;
                else {
                    const double *ch = pool.pool + 2ull * s_off[(t - 2) >> 1];
                    if (t & 1)
                        x = uq, y = ch, len = n1;
                    else
                        x = vq, y = ch + n1, len = n2;
                }
                double sum = 0.;
                for (int i = ln; i < len; i += 32)
                    sum = fma(x[i], y[i], sum);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1)
                    sum += __shfl_xor_sync(0xffffffffu, sum, d);
                if (ln == 0)
                    s_dot[t] = sum;
            }
            team_sync<TS>();
            const double aux_fast = fabs(s_dot[0]) * fabs(s_dot[1]);
            double cross          = 0.;
            for (int j = 0; j < q - 1; j++)
                cross += s_dot[2 + 2 * j] * s_dot[3 + 2 * j];
            frob_fast += aux_fast + 2. * cross;
            const double ratio = sqrt(aux_fast / frob_fast);
            constexpr double kGuard = 1e-6;
            const bool sure = dots_mode != 2 && isfinite(ratio) && fabs(ratio - epsilon) > kGuard * epsilon; // (uniform: every thread reads the same s_dot)
            if (sure)
                go_on = ratio > epsilon;
            else {
                aux = frob = 0.;
                for (int t1 = 1; t1 <= q; t1++) {
                    const double *ut = pool.pool + 2ull * s_off[t1 - 1];
                    team_sync<TS>();
                    in_order_estimator<TS>(t1, ut, ut + n1, n1, n2, pool.pool, s_off, s_dot, tid, aux, frob);
                }
                go_on = __dsqrt_rn(__ddiv_rn(aux, frob)) > epsilon;
            }
        }
    }
    team_sync<TS>();
    for (int j = tid; j < q; j += TS) // (q <= 0: nothing)
        pool.term_off[blk.term_base + j] = s_off[j];
    if (tid == 0)
        rank_out[bi] = q;
}

// ---- complex<double> kernel functions -----------------------------------------------------------------------------------------
// The same algorithm on complex coefficients, as the reference's template instantiated for std::complex<double> computes
// it (oracle/aca_oracle.c, oracle_sympartial_aca_z, pinned bit for bit against the live reference): |z| = std::abs (hypot),
// gamma = 1 / pivot by the compiler's complex division (libgcc __divdc3: Smith's formulas below; its rescaling of tiny /
// huge operands multiplies by powers of two and does not change the result in range), u2 *= gamma and the products of the dot
// products by the inline complex product (a c - b d, a d + b c), the dot products conjugate their first argument, the
// residual updates are OpenBLAS' zaxpy: y += ar x, then y += ai (i x), every product and every sum rounded (fma_axpy: the
// same two steps fused). A term is the chunk [uu_q (n1 complex) | vv_q (n2 complex)], re / im interleaved. The kernel
// functions without transcendental calls give the reference's factors bit for bit; Helmholtz differs from the host by the last
// ulps of sincos (and then stays within rounding of the reference: same pivots, same ranks).
struct Zc {
    double re, im;
};
__device__ __forceinline__ Zc zmul(Zc a, Zc b) { return Zc{__dsub_rn(__dmul_rn(a.re, b.re), __dmul_rn(a.im, b.im)), __dadd_rn(__dmul_rn(a.re, b.im), __dmul_rn(a.im, b.re))}; }
__device__ __forceinline__ Zc zinv(Zc z) { // (1 + 0 i) / z, libgcc __divdc3
    if (fabs(z.re) < fabs(z.im)) {
        const double ratio = __ddiv_rn(z.re, z.im), denom = __dadd_rn(__dmul_rn(z.re, ratio), z.im);
        return Zc{__ddiv_rn(ratio, denom), __ddiv_rn(-1., denom)};
    }
    const double ratio = __ddiv_rn(z.im, z.re), denom = __dadd_rn(__dmul_rn(z.im, ratio), z.re);
    return Zc{__ddiv_rn(1., denom), __ddiv_rn(-ratio, denom)};
}
template <bool FMA>
__device__ __forceinline__ Zc zaxpy1(Zc a, Zc x, Zc y) { // y + a x
    if (FMA)
        return Zc{fma(-a.im, x.im, fma(a.re, x.re, y.re)), fma(a.im, x.re, fma(a.re, x.im, y.im))};
    return Zc{__dsub_rn(__dadd_rn(y.re, __dmul_rn(a.re, x.re)), __dmul_rn(a.im, x.im)), __dadd_rn(__dadd_rn(y.im, __dmul_rn(a.re, x.im)), __dmul_rn(a.im, x.re))};
}
__device__ __forceinline__ Zc zload(const double *p, long long i) {
    const double2 v = *reinterpret_cast<const double2 *>(p + 2 * i);
    return Zc{v.x, v.y};
}
// sum += conj(x[i]) * y[i] in index order (wrapper_blas.hpp:152-157)
__device__ __forceinline__ Zc zdot_in_order(const double *x, const double *y, int len) {
    Zc s{0., 0.};
    for (int i = 0; i < len; i++) {
        const Zc a = zload(x, i), b = zload(y, i);
        s.re       = __dadd_rn(s.re, __dadd_rn(__dmul_rn(a.re, b.re), __dmul_rn(a.im, b.im)));  // a.re b.re - (-a.im) b.im
        s.im       = __dadd_rn(s.im, __dsub_rn(__dmul_rn(a.re, b.im), __dmul_rn(a.im, b.re)));  // a.re b.im + (-a.im) b.re
    }
    return s;
}

template <int KERNEL>
__device__ __forceinline__ Zc entry_z(bool swapped, const double *p1, const double *p2, int a, int b, double wavenumber) {
    const Value v = swapped ? kernel_value<KERNEL>(p2 + 3ll * b, p1 + 3ll * a, wavenumber) : kernel_value<KERNEL>(p1 + 3ll * a, p2 + 3ll * b, wavenumber);
    return Zc{v.re, v.im};
}

// In-order estimator, complex (sympartialACA.hpp:156-168): s_dot holds (re, im) pairs.
template <int TS>
__device__ __forceinline__ void in_order_estimator_z(int q, const double *uq, const double *vq, int n1, int n2, const double *pool, const uint32_t *s_off, double *s_dot, int tid, double &aux, double &frob) {
    const int ndots = 2 * q;
    for (int t = tid; t < ndots; t += TS) {
        Zc d;
        if (t == 0)
            d = zdot_in_order(uq, uq, n1);
        else if (t == 1)
            d = zdot_in_order(vq, vq, n2);
        else {
            const double *ch = pool + 2ull * s_off[(t - 2) >> 1];
            d                = (t & 1) ? zdot_in_order(uq, ch, n1) : zdot_in_order(vq, ch + 2ll * n1, n2);
        }
        s_dot[2 * t]     = d.re;
        s_dot[2 * t + 1] = d.im;
    }
    team_sync<TS>();
    aux = __dmul_rn(hypot(s_dot[0], s_dot[1]), hypot(s_dot[2], s_dot[3]));
    Zc frob_aux{0., 0.};
    for (int j = 0; j < q - 1; j++) {
        const Zc p   = zmul(Zc{s_dot[4 + 4 * j], s_dot[5 + 4 * j]}, Zc{s_dot[6 + 4 * j], s_dot[7 + 4 * j]}); // <u1, vv_j> <u2, uu_j>
        frob_aux.re  = __dadd_rn(frob_aux.re, p.re);
        frob_aux.im  = __dadd_rn(frob_aux.im, p.im);
    }
    frob = __dadd_rn(frob, __dadd_rn(aux, __dmul_rn(2., frob_aux.re)));
}

template <int TS, int KERNEL, bool FMA>
__global__ void __launch_bounds__(TS, TS == 512 ? 2 : (TS == 128 ? 6 : 20)) aca_kernel_z(const AcaBlock *__restrict__ blocks, long long first, const double *__restrict__ tp, const double *__restrict__ sp, double wavenumber, double epsilon,
                                                                                        AcaPool pool, int32_t *__restrict__ rank_out, int dots_mode) {
    constexpr int kCap = aca_max_rank(TS);
    __shared__ uint32_t s_off[kCap];
    __shared__ int s_piv1[kCap], s_piv2[kCap];
    __shared__ double s_dot[2 * (2 * kCap + 2)];
    __shared__ double s_rv[TS / 32];
    __shared__ int s_ri[TS / 32];
    __shared__ unsigned long long s_chunk;

    const long long bi = first + blockIdx.x;
    const AcaBlock blk = blocks[bi];
    const int tid      = threadIdx.x;
    const bool swapped = blk.swapped != 0;
    const int n1 = swapped ? blk.n : blk.m, n2 = swapped ? blk.m : blk.n;
    const double *p1 = swapped ? sp + 3ll * blk.lcol : tp + 3ll * blk.lrow;
    const double *p2 = swapped ? tp + 3ll * blk.lrow : sp + 3ll * blk.lcol;

    int q = 0, I1 = 0, I2 = 0, nv1 = 0, nv2 = 0;
    double frob = 0., aux = 0., frob_fast = 0.;
    bool go_on = true;
    while (go_on) {
        q += 1;
        if (static_cast<long long>(q) * (static_cast<long long>(n1) + n2) > static_cast<long long>(n1) * n2) {
            q = kAcaFailed;
            break;
        }
        if (q > static_cast<int>(blk.term_cap)) {
            q = kAcaRankCap;
            break;
        }
        const unsigned long long len = 2ull * (static_cast<unsigned long long>(n1) + n2); // doubles
        team_sync<TS>();
        if (tid == 0) {
            const unsigned long long off = atomicAdd(pool.cursor, len);
            s_chunk                      = off + len <= pool.capacity ? off : ~0ull;
            if (off + len <= pool.capacity)
                s_off[q - 1] = static_cast<uint32_t>(off >> 1);
        }
        team_sync<TS>();
        const unsigned long long off = s_chunk;
        if (off == ~0ull) {
            q = kAcaPoolOverflow;
            break;
        }
        double *uq = pool.pool + off, *vq = uq + 2ll * n1;

        double best = 0.;
        int idx     = -1;
        for (int b = tid; b < n2; b += TS) {
            Zc v = entry_z<KERNEL>(swapped, p1, p2, I1, b, wavenumber);
            for (int j = 0; j < q - 1; j++) {
                const double *ch = pool.pool + 2ull * s_off[j];
                const Zc c       = zload(ch, I1);
                v                = zaxpy1<FMA>(Zc{-c.re, -c.im}, zload(ch, static_cast<long long>(n1) + b), v);
            }
            *reinterpret_cast<double2 *>(vq + 2ll * b) = make_double2(v.re, v.im);
            bool seen = false;
            for (int t = 0; t < nv2; t++)
                seen |= s_piv2[t] == b;
            if (!seen) {
                const double tmp = hypot(v.re, v.im);
                if (tmp >= best) {
                    best = tmp;
                    idx  = b;
                }
            }
        }
        idx = team_argmax<TS>(best, idx, s_rv, s_ri);
        if (idx >= 0)
            I2 = idx;
        if (tid == 0)
            s_piv1[nv1] = I1;
        nv1++;
        team_sync<TS>();
        const Zc pivot = zload(vq, I2);
        if (!(hypot(pivot.re, pivot.im) > 1e-15)) {
            q -= 1;
            if (q == 0)
                q = kAcaFailed;
            break;
        }
        const Zc gamma = zinv(pivot);

        best = 0.;
        idx  = -1;
        for (int a = tid; a < n1; a += TS) {
            Zc v = entry_z<KERNEL>(swapped, p1, p2, a, I2, wavenumber);
            for (int j = 0; j < q - 1; j++) {
                const double *ch = pool.pool + 2ull * s_off[j];
                const Zc c       = zload(ch, static_cast<long long>(n1) + I2);
                v                = zaxpy1<FMA>(Zc{-c.re, -c.im}, zload(ch, a), v);
            }
            v = zmul(v, gamma);
            *reinterpret_cast<double2 *>(uq + 2ll * a) = make_double2(v.re, v.im);
            bool seen = false;
            for (int t = 0; t < nv1; t++)
                seen |= s_piv1[t] == a;
            if (!seen) {
                const double tmp = hypot(v.re, v.im);
                if (tmp >= best) {
                    best = tmp;
                    idx  = a;
                }
            }
        }
        idx = team_argmax<TS>(best, idx, s_rv, s_ri);
        if (idx >= 0)
            I1 = idx;
        if (tid == 0)
            s_piv2[nv2] = I2;
        nv2++;
        team_sync<TS>();

        if (TS == 32 || dots_mode == 1) {
            in_order_estimator_z<TS>(q, uq, vq, n1, n2, pool.pool, s_off, s_dot, tid, aux, frob);
            go_on = __dsqrt_rn(__ddiv_rn(aux, frob)) > epsilon;
        } else { // the guarded warp-level dot products of the real kernel, complex
            constexpr int W = TS / 32;
            const int w = tid >> 5, ln = tid & 31;
            const int ndots = 2 * q;
            for (int t = w; t < ndots; t += W) {
                const double *x, *y;
                int len;
                if (t == 0)
                    x = uq, y = uq, len = n1;
                else if (t == 1)
                    x = vq, y = vq, len = n2;
                else {
                    const double *ch = pool.pool + 2ull * s_off[(t - 2) >> 1];
                    if (t & 1)
                        x = uq, y = ch, len = n1;
                    else
                        x = vq, y = ch + 2ll * n1, len = n2;
                }
                double sr = 0., si = 0.;
                for (int i = ln; i < len; i += 32) {
                    const Zc a = zload(x, i), b = zload(y, i);
                    sr         = fma(a.re, b.re, fma(a.im, b.im, sr));
                    si         = fma(a.re, b.im, fma(-a.im, b.re, si));
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    sr += __shfl_xor_sync(0xffffffffu, sr, d);
                    si += __shfl_xor_sync(0xffffffffu, si, d);
                }
                if (ln == 0) {
                    s_dot[2 * t]     = sr;
                    s_dot[2 * t + 1] = si;
                }
            }
            team_sync<TS>();
            const double aux_fast = hypot(s_dot[0], s_dot[1]) * hypot(s_dot[2], s_dot[3]);
            double cross          = 0.;
            for (int j = 0; j < q - 1; j++)
                cross += s_dot[4 + 4 * j] * s_dot[6 + 4 * j] - s_dot[5 + 4 * j] * s_dot[7 + 4 * j];
            frob_fast += aux_fast + 2. * cross;
            const double ratio = sqrt(aux_fast / frob_fast);
            constexpr double kGuard = 1e-6;
            const bool sure = dots_mode != 2 && isfinite(ratio) && fabs(ratio - epsilon) > kGuard * epsilon;
            if (sure)
                go_on = ratio > epsilon;
            else {
                aux = frob = 0.;
                for (int t1 = 1; t1 <= q; t1++) {
                    const double *ut = pool.pool + 2ull * s_off[t1 - 1];
                    team_sync<TS>();
                    in_order_estimator_z<TS>(t1, ut, ut + 2ll * n1, n1, n2, pool.pool, s_off, s_dot, tid, aux, frob);
                }
                go_on = __dsqrt_rn(__ddiv_rn(aux, frob)) > epsilon;
            }
        }
    }
    team_sync<TS>();
    for (int j = tid; j < q; j += TS)
        pool.term_off[blk.term_base + j] = s_off[j];
    if (tid == 0)
        rank_out[bi] = q;
}

// One warp per unit of a low-rank leaf whose factors live in the pool: panel(i, k) = term (k0 + k), entry (p0 + i) of the
// side's vector — U column k = uu_k (vv_k if the block's dimensions were swapped), V row k = the other one.
template <bool CPLX>
__global__ void scatter_lowrank_kernel(const DenseTask *tasks, long long n_tasks, int side, unsigned char *stream, const AcaLeaf *leaves, const double *pool, const uint32_t *term_off) {
    const long long t = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tasks)
        return;
    const DenseTask task = tasks[t];
    const AcaLeaf lf     = leaves[task.lrow];
    const int lane       = threadIdx.x & 31;
    double *out          = reinterpret_cast<double *>(stream + task.byte_off);
    const bool second    = (side == 0) == (lf.swapped != 0); // the vv half of the chunk
    const size_t delta   = (second ? lf.n1 : 0u) + static_cast<size_t>(task.p0);
    const int total      = static_cast<int>(task.h) * task.w;
    for (int e = lane; e < total; e += 32) {
        const int k = e / task.h, i = e - k * task.h;
        const unsigned long long src = 2ull * term_off[lf.term_base + task.k0 + k];
        if (CPLX)
            reinterpret_cast<double2 *>(out)[static_cast<size_t>(k) * task.ld + i] = *reinterpret_cast<const double2 *>(pool + src + 2ull * (delta + i));
        else
            out[static_cast<size_t>(k) * task.ld + i] = pool[src + delta + i];
    }
}

template <int TS, int KERNEL>
cudaError_t launch_team(const AcaBlock *blocks, long long first, long long count, const double *tp, const double *sp, double wavenumber, double epsilon, int fma_axpy, AcaPool pool, int32_t *rank, int dots_mode, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>(count);
    if (fma_axpy)
        aca_kernel<TS, KERNEL, true><<<grid, TS, 0, st>>>(blocks, first, tp, sp, wavenumber, epsilon, pool, rank, dots_mode);
    else
        aca_kernel<TS, KERNEL, false><<<grid, TS, 0, st>>>(blocks, first, tp, sp, wavenumber, epsilon, pool, rank, dots_mode);
    return cudaGetLastError();
}

template <int TS, int KERNEL>
cudaError_t launch_team_z(const AcaBlock *blocks, long long first, long long count, const double *tp, const double *sp, double wavenumber, double epsilon, int fma_axpy, AcaPool pool, int32_t *rank, int dots_mode, cudaStream_t st) {
    const unsigned grid = static_cast<unsigned>(count);
    if (fma_axpy)
        aca_kernel_z<TS, KERNEL, true><<<grid, TS, 0, st>>>(blocks, first, tp, sp, wavenumber, epsilon, pool, rank, dots_mode);
    else
        aca_kernel_z<TS, KERNEL, false><<<grid, TS, 0, st>>>(blocks, first, tp, sp, wavenumber, epsilon, pool, rank, dots_mode);
    return cudaGetLastError();
}

template <int KERNEL>
cudaError_t launch_kernel_z(int team, const AcaBlock *blocks, long long first, long long count, const double *tp, const double *sp, double wavenumber, double epsilon, int fma_axpy, AcaPool pool, int32_t *rank, int dots_mode, cudaStream_t st) {
    switch (team) {
    case 32: return launch_team_z<32, KERNEL>(blocks, first, count, tp, sp, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case 128: return launch_team_z<128, KERNEL>(blocks, first, count, tp, sp, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case 512: return launch_team_z<512, KERNEL>(blocks, first, count, tp, sp, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    default: return cudaErrorInvalidValue;
    }
}

template <int KERNEL>
cudaError_t launch_kernel(int team, const AcaBlock *blocks, long long first, long long count, const double *tp, const double *sp, double wavenumber, double epsilon, int fma_axpy, AcaPool pool, int32_t *rank, int dots_mode, cudaStream_t st) {
    switch (team) {
    case 32: return launch_team<32, KERNEL>(blocks, first, count, tp, sp, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case 128: return launch_team<128, KERNEL>(blocks, first, count, tp, sp, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case 512: return launch_team<512, KERNEL>(blocks, first, count, tp, sp, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace

cudaError_t launch_aca(int kernel, int team, const AcaBlock *blocks, long long first, long long count, const double *target_points, const double *source_points, double wavenumber, double epsilon, int fma_axpy, AcaPool pool,
                       int32_t *rank, int dots_mode, cudaStream_t st) {
    if (count <= 0)
        return cudaSuccess;
    if (count > 0x7fffffffll)
        return cudaErrorInvalidValue;
    switch (kernel) {
    case HTB_KERNEL_COMPLEX_REG: return launch_kernel_z<HTB_KERNEL_COMPLEX_REG>(team, blocks, first, count, target_points, source_points, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case HTB_KERNEL_HERMITIAN_REG: return launch_kernel_z<HTB_KERNEL_HERMITIAN_REG>(team, blocks, first, count, target_points, source_points, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case HTB_KERNEL_HELMHOLTZ: return launch_kernel_z<HTB_KERNEL_HELMHOLTZ>(team, blocks, first, count, target_points, source_points, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case HTB_KERNEL_COMPLEX: return launch_kernel_z<HTB_KERNEL_COMPLEX>(team, blocks, first, count, target_points, source_points, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case HTB_KERNEL_LAPLACE: return launch_kernel<HTB_KERNEL_LAPLACE>(team, blocks, first, count, target_points, source_points, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    case HTB_KERNEL_LAPLACE_REG: return launch_kernel<HTB_KERNEL_LAPLACE_REG>(team, blocks, first, count, target_points, source_points, wavenumber, epsilon, fma_axpy, pool, rank, dots_mode, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_scatter_lowrank(bool complex_coefficients, const DenseTask *tasks, long long n_tasks, int side, unsigned char *stream, const AcaLeaf *leaves, const double *pool, const uint32_t *term_off, cudaStream_t st) {
    if (n_tasks == 0)
        return cudaSuccess;
    const int threads   = 256;
    const unsigned grid = static_cast<unsigned>((n_tasks * 32 + threads - 1) / threads);
    if (complex_coefficients)
        scatter_lowrank_kernel<true><<<grid, threads, 0, st>>>(tasks, n_tasks, side, stream, leaves, pool, term_off);
    else
        scatter_lowrank_kernel<false><<<grid, threads, 0, st>>>(tasks, n_tasks, side, stream, leaves, pool, term_off);
    return cudaGetLastError();
}

} // namespace htb
