// duo.cuh -- Fq2 arithmetic split over TWO adjacent lanes, for the line-schedule kernel.
//
// The thread-per-pairing line kernel has too few warps at the headline batch (2^14 pairings = 512 warps for 592 SM
// sub-partitions).  Here lanes (2j, 2j+1) own pairing j and lane h keeps ONLY component h ("its own" component) of
// every Fq2 value:
//   * linear operations (add, sub, neg, halve) are one Fq operation per lane -- no redundant work;
//   * for a product / square / xi-multiple a lane fetches the other component of the operands from its partner through
//     a small shared-memory area (D::partner*), computes component h of the result (two Fq products and ONE Montgomery
//     reduction for a product) and keeps it: results need no exchange.
// Round 2 measured that this kernel's time is its instruction count (weighted by pipe cost), not its dependency chain:
// the first lane-pair form (both lanes hold whole Fq2 values, linear operations done twice, results joined with
// shuffles + selects) took 1.56 ms, a four-lane form of this file (two operations per round, twice the warps, linear
// operations still done twice per pairing) 1.52 ms, this form 1.28 ms (profiles/README.md runs 27, 28).
// Same formulas, same canonical values as fp2.cuh; reference src/fields/fq2.rs.
#pragma once
#include "fp2.cuh"

namespace bn {

// Duo context D: h() in {0,1} (the component this lane owns), small_reduce9(v, out) = quotient-estimate reduction of a
// 9-limb value < 16q (k*q table in shared memory on the device),
//   partner(v)              the partner lane's v
//   partner2(a, b, ao, bo)  the same for two values (one exchange)
// Below m = own component, o = the partner's: lane 0 holds (m, o) = (c0, c1), lane 1 holds (c1, c0).
template <class D>
BN_HD Fp duo_own(const D& d, const Fp2& a) {
    return d.h() ? a.c1 : a.c0;
}
template <class D>
BN_HD Fp2 duo_whole(const D& d, const Fp& m) {
    Fp o = d.partner(m);
    return d.h() ? Fp2{o, m} : Fp2{m, o};
}

// The three out-of-line cores take OWN components and fetch the partner's themselves (half the argument registers of a
// version that is handed both components).
// component h of a * b: lane 0: a0 b0 + a1 (q - b1); lane 1: a1 b0 + a0 b1.     reference src/fields/fq2.rs:139-155
template <class D>
BN_HD_NOINLINE Fp duo_mul(const D d, Fp ma, Fp mb) {
    const bool h = d.h() != 0;
    Fp oa, ob;
    d.partner2(ma, mb, oa, ob);
    Fp y0 = fp_select(h, ob, mb);
    Fp y1 = fp_select(h, mb, fp_neg_lazy<MQ>(ob));
    return fp_mul2<MQ>(ma, y0, oa, y1);  // operands <= q: the sum is < 2 q^2
}
// component h of a^2: lane 0: (a0 + a1)(a0 + (q - a1)); lane 1: (2 a1) a0.       reference src/fields/fq2.rs:112-123
template <class D>
BN_HD_NOINLINE Fp duo_sqr(const D d, Fp m) {
    const bool h = d.h() != 0;
    Fp o = d.partner(m);
    Fp x = fp_add_raw(m, fp_select(h, m, o));
    Fp y = fp_select(h, o, fp_add_raw(m, fp_neg_lazy<MQ>(o)));
    return fp_mul<MQ>(x, y);  // x, y < 2q: x y < 4 q^2 < 2^256 q
}
// component h of xi * a: lane 0: 9 a0 - a1; lane 1: 9 a1 + a0.                   reference src/fields/fq2.rs:70-72
template <class D>
BN_HD_NOINLINE Fp duo_mul_xi(const D d, Fp m) {
    Fp o = d.partner(m);
    Fp addend = fp_select(d.h() != 0, o, fp_neg_lazy<MQ>(o));
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
    d.small_reduce9(v, r.v);
    return r;
}
// a * k, k in Fq: no exchange at all.                                             reference src/fields/fq2.rs:63-68
BN_HD Fp duo_mul_fp(const Fp& a, const Fp& k) { return fp_mul_ni<MQ>(a, k); }

}  // namespace bn
