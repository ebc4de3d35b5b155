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
//
// One Miller iteration: f <- f^2 (if do_sqr), then f <- f * line(t), ..., f * line(t + nlines - 1).
// Generic squaring (reference src/fields/fq12.rs:275-282) = 21 distinct products in 4 lock-step rounds; product with the
// sparse line l0 + l3 w^3 + l4 w^4 (reference mul_by_024, src/fields/fq12.rs:107-176) = 3 rounds.  Both run through ONE
// inlined multiply-accumulate inside this single out-of-line function (instruction-cache footprint of the Miller phase).
template <class Ctx, class LineSrc>
BN_HD_NOINLINE Fp2 hx_miller_iter(const Ctx c, const LineSrc src, Fp2 f, int t, int nlines, int do_sqr) {
    const int k = c.k();
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int op = do_sqr ? 0 : 1; op <= nlines; op++) {
        Fp2 l0 = fp2_zero(), l3k = l0, l4k = l0;
        c.sync();
        c.put(0, f);
        if (op == 0) {
            Fp2 xa = c.mul_xi(f);
            c.put(1, xa);
            c.put(2, fp2_dbl(fp2_select(k >= 4, xa, f)));  // doubled operand: 2 a_k on lanes 0..3, 2 xi a_k on lanes 4,5
        } else {
            src.get(t + op - 1, k, l0, l3k, l4k);
        }
        c.sync();
        AccK acc;
        acck_init(acc);
        const int nr = op == 0 ? 4 : 3;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int r = 0; r < nr; r++) {
            Fp2 x, y;
            if (op == 0) {
                // x sources / y sources per lane (nibble k), per round:
                //  r0: (2xi a5) a1 | (2a0) a1 | (2a0) a2 | (2a0) a3 | (2a0) a4 | (2a0) a5
                //  r1: (2xi a4) a2 | (2xi a5) a2 | (2xi a5) a3 | (2a1) a2 | (2a1) a3 | (2a1) a4
                //  r2: a0 a0 | (2xi a4) a3 | a1 a1 | (2xi a4) a5 | a2 a2 | (2a3) a2
                //  r3: xi a3 a3 | - | xi a4 a4 | - | xi a5 a5 | -
                const uint32_t xs = r == 0 ? 0x000005u : r == 1 ? 0x111554u : r == 2 ? 0x324140u : 0x050403u;
                const uint32_t ys = r == 0 ? 0x543211u : r == 1 ? 0x432322u : r == 2 ? 0x225130u : 0x050403u;
                const int xsrc = nib(xs, k);
                const int xslot = r < 2 ? 2 : (r == 2 ? ((xsrc == 3 || xsrc == 4) ? 2 : 0) : 1);
                x = c.get(xsrc, xslot);
                y = c.get(nib(ys, k), 0);
                y = fp2_select(r == 3 && (k & 1) != 0, fp2_zero(), y);
            } else {
                x = c.get(mod6(k + (r == 0 ? 0 : r == 1 ? 3 : 2)), 0);  // a_k, a_{k-3}, a_{k-4}
                y = fp2_select(r == 0, l0, fp2_select(r == 1, l3k, l4k));
            }
            mac_fp2(acc, x, y);
        }
        f = reduce2(acc);
    }
    return f;
}

// reference G2Precomp::miller_loop, src/groups/mod.rs:485-520
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
        const int nl = 1 + (int)((BN_ATE_NAF_NZ >> b) & 1ULL);
        f = hx_miller_iter(c, src, f, t, nl, 1);
        t += nl;
    }
#else
    if ((BN_ATE_BITS >> (BN_ATE_NBITS - 1)) & 1ULL) {
        f = hx_miller_iter(c, src, f, t, 1, 0);
        t += 1;
    }
    for (int b = BN_ATE_NBITS - 2; b >= 0; b--) {
        const int nl = 1 + (int)((BN_ATE_BITS >> b) & 1ULL);
        f = hx_miller_iter(c, src, f, t, nl, 1);
        t += nl;
    }
#endif
    return hx_miller_iter(c, src, f, t, 2, 0);  // the two Frobenius lines
}

// word offsets of the five Fq2 values inside a stored line (curve.cuh: struct Line)
#define BN_LINE_OFF_L0 0
#define BN_LINE_OFF_L3 16
#define BN_LINE_OFF_XL3 32
#define BN_LINE_OFF_L4 48
#define BN_LINE_OFF_XL4 64

}  // namespace bn
