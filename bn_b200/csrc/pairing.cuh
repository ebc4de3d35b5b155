// pairing.cuh -- the Miller-loop accumulation f <- f^2 * line over a hexad, fed by precomputed lines.
//
// reference G2Precomp::miller_loop, src/groups/mod.rs:485-520 (digits of 6u+2 below the leading one: square, multiply
// by the doubling line, and on non-zero digits by the addition line; then the two Frobenius lines).  The digit
// schedule (NAF by default, the reference's binary walk with BN_ATE_NAF=0) is the one ate_lines() emitted.
#pragma once
#include "curve.cuh"
#include "hexad.cuh"

namespace bn {

// LineSrc:  Handle acquire(t)        wait until line t is readable
//           Fp2 coef(h, i)            i = 0: l0, 1: l3k = (k < 3 ? xi*l3 : l3), 2: l4k = (k < 4 ? xi*l4 : l4) for this lane
//           void release(t)           line t has been consumed (the device ring refills its buffer with line t + 2)
template <class Ctx, class LineSrc>
BN_HD Fp2 hx_miller_loop(const Ctx& c, const LineSrc& src) {
    int t = 0;
    // first iteration: f = 1, so f^2 * line is the line itself: l0 + l3 w^3 + l4 w^4 (lanes 3 and 4 hold the
    // plain l3 / l4 variants, see LineSrc::coef)
    Fp2 f;
    {
        typename LineSrc::Handle h = src.acquire(t);
        const int k = c.k();
        f = src.coef(h, k == 3 ? 1 : (k == 4 ? 2 : 0));
        f = fp2_select(k == 0 || k == 3 || k == 4, f, fp2_zero());
        src.release(t++);
    }
    auto line_step = [&](Fp2 v) {
        typename LineSrc::Handle h = src.acquire(t);
        Fp2 r = hx_mul_line(c, v, src, h);
        src.release(t++);
        return r;
    };
#if BN_ATE_NAF
    // digit BN_ATE_NAF_DIGITS-1 (= 64) is zero: no addition after the first doubling
    for (int b = BN_ATE_NAF_DIGITS - 2; b >= 0; b--) {
        f = hx_sqr(c, f);
        f = line_step(f);
        if ((BN_ATE_NAF_NZ >> b) & 1ULL) f = line_step(f);
    }
#else
    if ((BN_ATE_BITS >> (BN_ATE_NBITS - 1)) & 1ULL) f = line_step(f);
    for (int b = BN_ATE_NBITS - 2; b >= 0; b--) {
        f = hx_sqr(c, f);
        f = line_step(f);
        if ((BN_ATE_BITS >> b) & 1ULL) f = line_step(f);
    }
#endif
    for (int e = 0; e < 2; e++) f = line_step(f);
    return f;
}

// word offsets of the five Fq2 values inside a stored line (curve.cuh: struct Line)
#define BN_LINE_OFF_L0 0
#define BN_LINE_OFF_L3 16
#define BN_LINE_OFF_XL3 32
#define BN_LINE_OFF_L4 48
#define BN_LINE_OFF_XL4 64

}  // namespace bn
