// pairing.cuh -- the Miller-loop accumulation f <- f^2 * line over a hexad, fed by precomputed lines.
//
// reference G2Precomp::miller_loop, src/groups/mod.rs:485-520 (bits of 6u+2 below the MSB: square, multiply
// by the doubling line, and on set bits by the addition line; then the two Frobenius lines).
#pragma once
#include "curve.cuh"
#include "hexad.cuh"

namespace bn {

// LineSrc::get(t, k, l0, l3k, l4k): fetch line t with the xi-variants lane k needs
//   l3k = (k < 3 ? xi*l3 : l3),  l4k = (k < 4 ? xi*l4 : l4)
template <class Ctx, class LineSrc>
BN_HD Fp2 hx_miller_loop(const Ctx& c, const LineSrc& src) {
    Fp2 f = hx_one(c);
    Fp2 l0, l3k, l4k;
    int t = 0;
    for (int b = BN_ATE_NBITS - 1; b >= 0; b--) {
        f = hx_sqr(c, f);
        src.get(t++, c.k(), l0, l3k, l4k);
        f = hx_mul_line(c, f, l0, l3k, l4k);
        if ((BN_ATE_BITS >> b) & 1ULL) {
            src.get(t++, c.k(), l0, l3k, l4k);
            f = hx_mul_line(c, f, l0, l3k, l4k);
        }
    }
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
