// htool_b200/csrc/kernel_functions.cuh — the built-in kernel functions (htb_generator_desc.kernel) evaluated on the device, shared
// by the dense-leaf generation (generate.cu) and the batched ACA (aca.cu).
//
// Built-in kernel functions = the analytic generators of the reference's test-suite
// (include/htool/testing/generator_test.hpp:155-205) plus the Helmholtz kernel of SURVEY.md 8d. Every operation is an
// explicitly rounded IEEE operation in the order the host generator performs it (no FMA contraction), so the real
// kernels are BIT-IDENTICAL to the host generators; the Helmholtz kernel differs by the last ulps of sin / cos.
#ifndef HTB_KERNEL_FUNCTIONS_CUH
#define HTB_KERNEL_FUNCTIONS_CUH

#include <htool_b200.h>

namespace htb {

struct Value {
    double re, im;
};

template <int KERNEL>
__device__ __forceinline__ Value kernel_value(const double *a, const double *b, double wavenumber) {
    const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
    const double r  = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    const double four_pi = 4 * 3.14159265358979323846; // 4 * M_PI, folded at compile time on the host as well
    const double fpr     = __dmul_rn(four_pi, r);
    if (KERNEL == HTB_KERNEL_LAPLACE) // 1 / (4 pi r), generator_test.hpp:155-161
        return Value{__ddiv_rn(1., fpr), 0.};
    if (KERNEL == HTB_KERNEL_LAPLACE_REG) // 1 / (1e-5 + 4 pi r), :180-187
        return Value{__ddiv_rn(1., __dadd_rn(1e-5, fpr)), 0.};
    if (KERNEL == HTB_KERNEL_COMPLEX) { // (1 + i) / (4 pi r), :163-170
        const double v = __ddiv_rn(1., fpr);
        return Value{v, v};
    }
    if (KERNEL == HTB_KERNEL_COMPLEX_REG) { // (1 + i) / (1e-5 + 4 pi r), :189-196
        const double v = __ddiv_rn(1., __dadd_rn(1e-5, fpr));
        return Value{v, v};
    }
    if (KERNEL == HTB_KERNEL_HERMITIAN_REG) { // (1 + sign(x_t - x_s) i) / (1e-5 + 4 pi r), :198-205
        const double d = __dadd_rn(1e-5, fpr);
        const double s = dx > 0 ? 1. : (dx < 0 ? -1. : 0.);
        return Value{__ddiv_rn(1., d), __ddiv_rn(s, d)};
    }
    // HTB_KERNEL_HELMHOLTZ: exp(i k r) / (4 pi r), finite diagonal
    if (r < 1e-12)
        return Value{__ddiv_rn(1., __dmul_rn(four_pi, 1e-3)), __ddiv_rn(wavenumber, four_pi)};
    double sn, cs;
    sincos(__dmul_rn(wavenumber, r), &sn, &cs);
    return Value{__ddiv_rn(cs, fpr), __ddiv_rn(sn, fpr)};
}

} // namespace htb
#endif
