// duo.cuh -- Fq2 arithmetic split over TWO adjacent lanes, for the line-schedule kernel.
//
// The thread-per-pairing line kernel is latency-bound at the headline batch (2^14 pairings = 512 warps for 592 SM
// sub-partitions).  Here lanes (2j, 2j+1) own one pairing: both lanes hold every Fq2 value in full, but an Fq2
// product / square / scaling is computed one OUTPUT COMPONENT per lane (lane h computes c_h: two Fq products and one
// Montgomery reduction instead of four and two), and the halves are swapped with 8 warp shuffles.  Cheap linear
// operations (add, sub, neg, halve) are done redundantly by both lanes.  Same formulas, same canonical results as
// fp2.cuh; reference src/fields/fq2.rs.
#pragma once
#include "fp2.cuh"

namespace bn {

// Duo context D: h() in {0,1} (which component this lane produces), swap(v) = the partner lane's v,
// small_reduce9(v, out) = quotient-estimate reduction of a 9-limb value < 16q (k*q table in shared memory on the device).
template <class D>
BN_HD Fp2 duo_join(const D& d, const Fp& mine) {
    Fp other = d.swap(mine);
    return d.h() ? Fp2{other, mine} : Fp2{mine, other};
}

// The half of each operation that one lane computes (component h of the result).  Shared with quad.cuh.
// reference src/fields/fq2.rs:139-155
BN_HD Fp duo_mul_half(bool h, const Fp2& a, const Fp2& b) {
    // lane 0: a0*b0 + a1*(q - b1) ; lane 1: a0*b1 + a1*b0
    Fp y0 = fp_select(h, b.c1, b.c0);
    Fp y1 = fp_select(h, b.c0, fp_neg_lazy<MQ>(b.c1));
    Wide t = wide_zero();
    wide_mac2(t, a.c0, y0, a.c1, y1);
    return mont_reduce<MQ, 2>(t);
}
// reference src/fields/fq2.rs:112-123
BN_HD Fp duo_sqr_half(bool h, const Fp2& a) {
    // lane 0: (a0 + a1)(a0 + (q - a1)) ; lane 1: 2 * a0*a1
    Fp x = fp_select(h, a.c0, fp_add_raw(a.c0, a.c1));
    Fp y = fp_select(h, a.c1, fp_add_raw(a.c0, fp_neg_lazy<MQ>(a.c1)));
    Wide t = wide_zero();
    wide_mac1(t, x, y);
    Wide t2 = t;
    wide_dbl(t2);
    BN_UNROLL
    for (int i = 0; i < 16; i++) t.w[i] = h ? t2.w[i] : t.w[i];
    return mont_reduce<MQ, 2>(t);
}
// xi * a, one component per lane: lane 0: 9 a0 - a1, lane 1: 9 a1 + a0.   reference src/fields/fq2.rs:70-72
template <class D>
BN_HD Fp duo_xi_half(const D& d, const Fp2& a) {
    const bool h = d.h() != 0;
    Fp x = fp_select(h, a.c1, a.c0);
    Fp addend = fp_select(h, a.c0, fp_neg_lazy<MQ>(a.c1));
    uint32_t v[9];
    v[0] = x.v[0] << 3;
    BN_UNROLL
    for (int i = 1; i < 8; i++) v[i] = (x.v[i] << 3) | (x.v[i - 1] >> 29);
    v[8] = x.v[7] >> 29;
    uint32_t c = addi8(v, x.v);
    v[8] += c;
    c = addi8(v, addend.v);
    v[8] += c;
    Fp r;
    d.small_reduce9(v, r.v);
    return r;
}

template <class D>
BN_HD_NOINLINE Fp2 duo_mul(const D d, Fp2 a, Fp2 b) {
    return duo_join(d, duo_mul_half(d.h() != 0, a, b));
}
template <class D>
BN_HD_NOINLINE Fp2 duo_sqr(const D d, Fp2 a) {
    return duo_join(d, duo_sqr_half(d.h() != 0, a));
}
// reference src/fields/fq2.rs:63-68
template <class D>
BN_HD_NOINLINE Fp2 duo_mul_fp(const D d, Fp2 a, Fp k) {
    return duo_join(d, fp_mul<MQ>(d.h() ? a.c1 : a.c0, k));
}
template <class D>
BN_HD_NOINLINE Fp2 duo_mul_xi(const D d, Fp2 a) {
    return duo_join(d, duo_xi_half(d, a));
}

}  // namespace bn
