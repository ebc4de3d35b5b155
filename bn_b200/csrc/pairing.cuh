// pairing.cuh -- the Miller-loop accumulation f <- f^2 * line over a hexad, fed by precomputed lines.
//
// reference G2Precomp::miller_loop, src/groups/mod.rs:485-520 (digits of 6u+2 below the leading one: square, multiply
// by the doubling line, and on non-zero digits by the addition line; then the two Frobenius lines).  The digit
// schedule (NAF by default, the reference's binary walk with BN_ATE_NAF=0) is the one ate_lines() emitted.
#pragma once
#include "curve.cuh"
#include "hexad.cuh"

namespace bn {

// LineSrc::get(t, k, l0, l3k, l4k): fetch line t with the xi-variants lane k needs
//   l3k = (k < 3 ? xi*l3 : l3),  l4k = (k < 4 ? xi*l4 : l4)
template <class Ctx, class LineSrc>
BN_HD Fp2 hx_miller_loop(const Ctx& c, const LineSrc& src) {
    Fp2 l0, l3k, l4k;
    int t = 0;
    // first iteration: f = 1, so f^2 * line is the line itself: l0 + l3 w^3 + l4 w^4 (lanes 3 and 4 hold the
    // plain l3 / l4 variants, see LineSrc::get)
    src.get(t++, c.k(), l0, l3k, l4k);
    Fp2 f = fp2_select(c.k() == 0, l0, fp2_select(c.k() == 3, l3k, fp2_select(c.k() == 4, l4k, fp2_zero())));
#if BN_ATE_NAF
    // digit BN_ATE_NAF_DIGITS-1 (= 64) is zero: no addition after the first doubling
    for (int b = BN_ATE_NAF_DIGITS - 2; b >= 0; b--) {
        f = hx_sqr(c, f);
        src.get(t++, c.k(), l0, l3k, l4k);
        f = hx_mul_line(c, f, l0, l3k, l4k);
        if ((BN_ATE_NAF_NZ >> b) & 1ULL) {
            src.get(t++, c.k(), l0, l3k, l4k);
            f = hx_mul_line(c, f, l0, l3k, l4k);
        }
    }
#else
    if ((BN_ATE_BITS >> (BN_ATE_NBITS - 1)) & 1ULL) {
        src.get(t++, c.k(), l0, l3k, l4k);
        f = hx_mul_line(c, f, l0, l3k, l4k);
    }
    for (int b = BN_ATE_NBITS - 2; b >= 0; b--) {
        f = hx_sqr(c, f);
        src.get(t++, c.k(), l0, l3k, l4k);
        f = hx_mul_line(c, f, l0, l3k, l4k);
        if ((BN_ATE_BITS >> b) & 1ULL) {
            src.get(t++, c.k(), l0, l3k, l4k);
            f = hx_mul_line(c, f, l0, l3k, l4k);
        }
    }
#endif
    for (int e = 0; e < 2; e++) {
        src.get(t++, c.k(), l0, l3k, l4k);
        f = hx_mul_line(c, f, l0, l3k, l4k);
    }
    return f;
}

// word offsets of the five Fq2 values inside a stored line (curve.cuh: struct Line)
#define BN_LINE_OFF_L0 0
#define BN_LINE_OFF_L3 16
#define BN_LINE_OFF_XL3 32
#define BN_LINE_OFF_L4 48
#define BN_LINE_OFF_XL4 64

}  // namespace bn
