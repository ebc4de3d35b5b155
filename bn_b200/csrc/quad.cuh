// quad.cuh -- TWO independent Fq2 operations at a time on FOUR adjacent lanes, for the line-schedule kernel.
//
// The lane-pair kernel (duo.cuh) is a dependent chain of ~600 k instructions per warp and has only 1024 warps at the
// headline batch (1.7 per scheduler): each warp issues one instruction every ~5 cycles and the IMAD pipe is half idle.
// The doubling and addition steps hold pairs of independent products (src/groups/mod.rs:592-634: y^2 and z^2, d^2 and
// e^2, ...), so here lanes 4j .. 4j+3 own pairing j and work in ROUNDS of two operations: lane 2s + h computes
// component h of the operation of side s (A = side 0, B = side 1).
//
// Component split: a lane keeps only component h ("its own" component) of every Fq2 value, so the linear operations
// between products (a third of the lane-pair kernel's instructions) cost one Fq operation per lane instead of two.
// Before a product a lane fetches the other component of its two operands from its partner (lane ^ 1); after it, the
// two sides swap results (lane ^ 2).  Both exchanges go through a small shared-memory area (Q::partner*, Q::sides).
// Twice the warps of duo.cuh and ~0.45x the instructions per warp; same formulas, same canonical values as fp2.cuh.
#pragma once
#include "duo.cuh"

namespace bn {

// Quad context Q: h() in {0,1} (component), s() in {0,1} (side), small_reduce9(v, out) as in duo.cuh,
//   partner(v)              lane ^ 1's v
//   partner2(a, b, ao, bo)  the same for two values
//   sides(mine, rA, rB)     rA = side 0's mine, rB = side 1's mine (same component)
// Below m = own component, o = the partner's: lane 0 holds (m, o) = (c0, c1), lane 1 holds (c1, c0).

template <class Q>
BN_HD Fp quad_own(const Q& q, const Fp2& a) {
    return q.h() ? a.c1 : a.c0;
}

// component h of a * b: lane 0: a0 b0 + a1 (q - b1); lane 1: a1 b0 + a0 b1.     reference src/fields/fq2.rs:139-155
template <class Q>
BN_HD_NOINLINE Fp quad_mul_core(const Q q, Fp ma, Fp oa, Fp mb, Fp ob) {
    const bool h = q.h() != 0;
    Fp y0 = fp_select(h, ob, mb);
    Fp y1 = fp_select(h, mb, fp_neg_lazy<MQ>(ob));
    Wide t = wide_zero();
    wide_mac2(t, ma, y0, oa, y1);
    return mont_reduce<MQ, 2>(t);
}
// component h of a^2: lane 0: (a0 + a1)(a0 + (q - a1)); lane 1: (2 a1) a0.       reference src/fields/fq2.rs:112-123
template <class Q>
BN_HD_NOINLINE Fp quad_sqr_core(const Q q, Fp m, Fp o) {
    const bool h = q.h() != 0;
    Fp x = fp_add_raw(m, fp_select(h, m, o));
    Fp y = fp_select(h, o, fp_add_raw(m, fp_neg_lazy<MQ>(o)));
    Wide t = wide_zero();
    wide_mac1(t, x, y);
    return mont_reduce<MQ, 2>(t);
}
// component h of xi * a: lane 0: 9 a0 - a1; lane 1: 9 a1 + a0.                   reference src/fields/fq2.rs:70-72
template <class Q>
BN_HD_NOINLINE Fp quad_xi_core(const Q q, Fp m, Fp o) {
    Fp addend = fp_select(q.h() != 0, o, fp_neg_lazy<MQ>(o));
    uint32_t v[9];
    v[0] = m.v[0] << 3;
    BN_UNROLL
    for (int i = 1; i < 8; i++) v[i] = (m.v[i] << 3) | (m.v[i - 1] >> 29);
    v[8] = m.v[7] >> 29;
    uint32_t c = addi8(v, m.v);
    v[8] += c;
    c = addi8(v, addend.v);
    v[8] += c;
    Fp r;
    q.small_reduce9(v, r.v);
    return r;
}

// One round.  Arguments and results are OWN components.  With Q::SIDES == 2 (four lanes) side A computes the first
// operation and side B the second; with Q::SIDES == 1 (a lane pair, no redundant work at all) the pair computes both.
// (rA, rB) = (aA * bA, aB * bB)
template <class Q>
BN_HD void quad_mul2(const Q& q, const Fp& aA, const Fp& bA, const Fp& aB, const Fp& bB, Fp& rA, Fp& rB) {
    Fp ao, bo;
    if constexpr (Q::SIDES == 2) {
        const bool s = q.s() != 0;
        Fp a = fp_select(s, aB, aA), b = fp_select(s, bB, bA);
        q.partner2(a, b, ao, bo);
        q.sides(quad_mul_core(q, a, ao, b, bo), rA, rB);
    } else {
        q.partner2(aA, bA, ao, bo);
        Fp r = quad_mul_core(q, aA, ao, bA, bo);
        q.partner2(aB, bB, ao, bo);
        rB = quad_mul_core(q, aB, ao, bB, bo);
        rA = r;
    }
}
// a * b alone (four lanes: both sides compute it)
template <class Q>
BN_HD Fp quad_mul1(const Q& q, const Fp& a, const Fp& b) {
    Fp ao, bo;
    q.partner2(a, b, ao, bo);
    return quad_mul_core(q, a, ao, b, bo);
}
// (rA, rB) = (aA^2, aB^2)
template <class Q>
BN_HD void quad_sqr2(const Q& q, const Fp& aA, const Fp& aB, Fp& rA, Fp& rB) {
    if constexpr (Q::SIDES == 2) {
        Fp a = fp_select(q.s() != 0, aB, aA);
        q.sides(quad_sqr_core(q, a, q.partner(a)), rA, rB);
    } else {
        Fp ao, bo;
        q.partner2(aA, aB, ao, bo);
        Fp r = quad_sqr_core(q, aA, ao);
        rB = quad_sqr_core(q, aB, bo);
        rA = r;
    }
}
template <class Q>
BN_HD Fp quad_sqr1(const Q& q, const Fp& a) {
    return quad_sqr_core(q, a, q.partner(a));
}
// The tail of a line: l3 = vw * py, l4 = vv * px (Fq scalings, src/fields/fq2.rs:63-68) and the xi-multiples the Miller
// kernel consumes.  Four lanes: side 0 produces (l3, xl3), side 1 (l4, xl4) -- only those fields are meaningful there.
struct LineH {
    Fp l0, l3, xl3, l4, xl4;
};
template <class Q>
BN_HD LineH quad_finish_line(const Q& q, const Fp& ell_0_pre, const Fp& vw, const Fp& vv, const Fp& px, const Fp& py) {
    LineH L;
    Fp ao, bo;
    if constexpr (Q::SIDES == 2) {
        const bool s = q.s() != 0;
        Fp l34 = fp_mul_ni<MQ>(fp_select(s, vv, vw), fp_select(s, px, py));
        q.partner2(ell_0_pre, l34, ao, bo);
        L.l0 = quad_xi_core(q, ell_0_pre, ao);
        Fp x = quad_xi_core(q, l34, bo);
        L.l3 = L.l4 = l34;
        L.xl3 = L.xl4 = x;
    } else {
        L.l3 = fp_mul_ni<MQ>(vw, py);
        L.l4 = fp_mul_ni<MQ>(vv, px);
        q.partner2(L.l3, L.l4, ao, bo);
        L.xl3 = quad_xi_core(q, L.l3, ao);
        L.xl4 = quad_xi_core(q, L.l4, bo);
        L.l0 = quad_xi_core(q, ell_0_pre, q.partner(ell_0_pre));
    }
    return L;
}

}  // namespace bn
