// hexad.cuh -- Fq12 / Gt arithmetic spread over SIX lanes of a warp ("hexad"), live values in registers.
//
// Representation.  The reference's tower Fq12 = Fq6[w]/(w^2 - v), Fq6 = Fq2[v]/(v^3 - xi)
// (src/fields/fq12.rs:26-31, src/fields/fq6.rs:42-48) is the same ring as Fq2[w]/(w^6 - xi):
//     c0.c0 + c1.c0 w + c0.c1 w^2 + c1.c1 w^3 + c0.c2 w^4 + c1.c2 w^5 .
// Lane k of a hexad (k = lane % 6; five hexads per warp, lanes 30/31 idle) holds the Fq2 coefficient g_k
// of w^k of EVERY live Fq12 value, so a whole Gt costs 16 registers per lane; values that stay idle for a whole
// exponentiation are parked in lane-private shared memory (c.park / c.unpark), everything else is in registers.
//
// Multiplication is the length-6 negacyclic-style convolution
//     c_k = sum_{i+j=k} a_i b_j + xi * sum_{i+j=k+6} a_i b_j
// evaluated one Fq2 product per lane per step; operands travel between lanes through shared-memory slots (below),
// and the products of a lane are summed as 512-bit integers and reduced once
// (2 Montgomery reductions per lane per Fq12 operation).  Results are canonical, hence bit-identical to the
// reference's Karatsuba tower (src/fields/fq12.rs:275-307, src/fields/fq6.rs:113-158) which computes the same
// ring element.
//
// Operand exchange: each lane owns three 64-byte slots in shared memory (c.put / c.get, ordered by c.sync ==
// __syncwarp: a hexad never spans warps).  A lane publishes a_k, xi*a_k and b_k once per operation and every round
// reads the operand it needs straight from the owner's slot -- the receiver picks the xi / doubled variant by
// address, so the round loop has no shuffles and no selects, and a, xi*a, b do not occupy registers while the
// 512-bit accumulators are live.
//
// All functions are collective over the hexad: every lane calls them with its own coefficient.
// `Ctx` supplies k(), put(slot, value), get(source lane index within the hexad, slot), get_or_zero(cond, lane, slot),
// small_reduce(lazy 9-limb integer < 16q), sync(), mul_xi(Fq2),
// park(slot, value) / unpark(slot) (lane-private storage outside the register file) and
// inv(Fq element, identical in the six lanes); kernels.cu binds them to shared memory + __syncwarp and a block-wide batched
// inversion, tests/host_emu binds them to a barrier exchange between six host threads and a Fermat inversion.
#pragma once
#include "fp2.cuh"
#ifndef BN_F52
#define BN_F52 1   // 1: Fq2 multiply-accumulate + reduction on the FP64 pipe (f52.cuh); 0: IMAD.WIDE accumulators (A/B)
#endif
#include "f52.cuh"
// per-operation A/B switches of the product path (all default to BN_F52)
#ifndef BN_F52_LINE
#define BN_F52_LINE BN_F52
#endif
#ifndef BN_F52_CYC
#define BN_F52_CYC BN_F52
#endif
#ifndef BN_F52_SQR
#define BN_F52_SQR BN_F52
#endif

namespace bn {

#if BN_F52
// Operand exchange in the FP64 form: a lane publishes an Fq2 value as three 5 x 52-bit factors (c0, c1, c0 + c1; D5x3,
// one 144-byte slot) with c.put52, and the receivers name a published operand by a Ctx::Ref (c.ref(lane, slot) /
// c.ref_or_zero(cond, lane, slot)) and load one factor at a time (c.ld5(ref, which)), so that only 2 x 5 doubles are live
// next to the column accumulators.
#ifndef BN_F52_PREFETCH
#define BN_F52_PREFETCH 0   // 1: issue the loads of the next factor pair before the current product's arithmetic (+20 registers)
#endif
template <class Ctx>
BN_HD void mac52_rr(const Ctx& c, Acc52& A, typename Ctx::Ref x, typename Ctx::Ref y) {
#if BN_F52_PREFETCH
    D5 a0 = c.ld5(x, 0), b0 = c.ld5(y, 0);
    D5 a1 = c.ld5(x, 1), b1 = c.ld5(y, 1);
    f52_mac(A.s0, a0, b0);
    a0 = c.ld5(x, 2);
    b0 = c.ld5(y, 2);
    f52_mac(A.s1, a1, b1);
    f52_mac(A.s2, a0, b0);
#else
    f52_mac(A.s0, c.ld5(x, 0), c.ld5(y, 0));
    f52_mac(A.s1, c.ld5(x, 1), c.ld5(y, 1));
    f52_mac(A.s2, c.ld5(x, 2), c.ld5(y, 2));
#endif
}
template <class Ctx>
BN_HD void mac52_rv(const Ctx& c, Acc52& A, typename Ctx::Ref x, const D5x3& y) {
    f52_mac(A.s0, c.ld5(x, 0), y.c0);
    f52_mac(A.s1, c.ld5(x, 1), y.c1);
    f52_mac(A.s2, c.ld5(x, 2), y.cs);
}
#endif

// ------------------------------------------------------------------------------------------------
// Lazy Fq2 multiply-accumulate on two merged 512-bit accumulators (arithmetic mod 2^512).
//   a0 += x0 y0 - x1 y1          a1 += (x0+x1)(y0+y1) - x0 y0 - x1 y1        (Karatsuba: 3 x 64 IMAD.WIDE)
// a0 starts at 6 q^2 (a multiple of q) so its true value never ends negative; a1's true value is
// x0 y1 + x1 y0 >= 0, intermediate wrap-around is harmless.  Operands may be canonical or "lazy" (< 2q per
// component, sums < 4q): every product stays < 16 q^2 < 2^512 and the final values stay < 12 q^2, which one
// Montgomery reduction plus two conditional subtractions brings back to canonical form.
// IMAD.WIDE is the scarce resource on B200 (quarter-rate), the ~170 IADD3 per step ride on the ALU pipe.
// ------------------------------------------------------------------------------------------------
// Two accumulator layouts (both measured, profiles/README.md):
//   AccK2  two merged 512-bit accumulators; every product is merged and added/subtracted as it is produced
//          (~190 IADD3 per round, 32 accumulator registers)
//   AccK3  three carry-save accumulators P0 = sum x0 y0, P1 = sum x1 y1, P2 = sum (x0+x1)(y0+y1); the round loop is
//          almost pure IMAD.WIDE (one IADD3.X per 4-IMAD chain) and the Karatsuba recombination happens once per
//          reduction (~64 IADD3 per round + ~160 per operation, 120 accumulator registers)
// The choice is made per operation (BN_ACC_MUL / _SQR / _LINE / _CYC = 2, 3 or 4 (AccK2S, below)).
struct AccK2 {
    Wide a0, a1;
};
BN_HD void acck_init(AccK2& A) {
    BN_UNROLL
    for (int i = 0; i < 16; i++) {
        A.a0.w[i] = WIDE_6Q2_f(i);
        A.a1.w[i] = 0;
    }
}
BN_HD void mac_fp2(AccK2& A, const Fp2& x, const Fp2& y) {
    Wide T;
    wide_mul(T, x.c0, y.c0);
    add16(A.a0.w, T.w);
    sub16(A.a1.w, T.w);
    wide_mul(T, x.c1, y.c1);
    sub16(A.a0.w, T.w);
    sub16(A.a1.w, T.w);
    wide_mul(T, fp_add_raw(x.c0, x.c1), fp_add_raw(y.c0, y.c1));
    add16(A.a1.w, T.w);
}
BN_HD void acck_finish(const AccK2& A, Wide& a0, Wide& a1) {
    a0 = A.a0;
    a1 = A.a1;
}
struct AccK3 {
    AccEO p0, p1, p2;
};
BN_HD void acck_init(AccK3& A) {
    acc_zero(A.p0);
    acc_zero(A.p1);
    acc_zero(A.p2);
}
BN_HD void mac_fp2(AccK3& A, const Fp2& x, const Fp2& y) {
    acc_mac(A.p0, x.c0, y.c0);
    acc_mac(A.p1, x.c1, y.c1);
    acc_mac(A.p2, fp_add_raw(x.c0, x.c1), fp_add_raw(y.c0, y.c1));
}
BN_HD void acck_finish(const AccK3& A, Wide& a0, Wide& a1) {
    Wide t0 = acc_merge(A.p0), t1 = acc_merge(A.p1);
    a1 = acc_merge(A.p2);
    sub16(a1.w, t0.w);
    sub16(a1.w, t1.w);
    BN_UNROLL
    for (int i = 0; i < 16; i++) a0.w[i] = WIDE_6Q2_f(i);
    add16(a0.w, t0.w);
    sub16(a0.w, t1.w);
}
// AccK2S: three merged 512-bit sums S0 = sum x0 y0, S1 = sum x1 y1, S2 = sum (x0+x1)(y0+y1) (mod 2^512; S2 may wrap, the
// recombination below is mod 2^512 too and its true value is in range).  3 x 16 accumulate IADD3 per round instead of
// AccK2's 5 x 16, one 4 x 16 recombination per operation; pays off from two rounds up (dense product, squaring).
struct AccK2S {
    Wide s0, s1, s2;
};
BN_HD void acck_init(AccK2S& A) {
    A.s0 = wide_zero();
    A.s1 = wide_zero();
    A.s2 = wide_zero();
}
BN_HD void mac_fp2(AccK2S& A, const Fp2& x, const Fp2& y) {
    wide_mac1(A.s0, x.c0, y.c0);
    wide_mac1(A.s1, x.c1, y.c1);
    wide_mac1(A.s2, fp_add_raw(x.c0, x.c1), fp_add_raw(y.c0, y.c1));
}
BN_HD void acck_finish(const AccK2S& A, Wide& a0, Wide& a1) {
    BN_UNROLL
    for (int i = 0; i < 16; i++) a0.w[i] = WIDE_6Q2_f(i);
    add16(a0.w, A.s0.w);
    sub16(a0.w, A.s1.w);
    a1 = A.s2;
    sub16(a1.w, A.s0.w);
    sub16(a1.w, A.s1.w);
}
template <int N>
struct AccSel {
    typedef AccK2 type;
};
template <>
struct AccSel<3> {
    typedef AccK3 type;
};
template <>
struct AccSel<4> {
    typedef AccK2S type;
};
// run 32 (dynamic instructions per call, tools/dyn_counts.py): AccK2S against the previous choice -- dense product 2893
// vs 3077 (AccK2), squaring 2298 vs 2391 (AccK2), line product 1603 vs 1728 (AccK3; k_miller also drops to 158 registers).
#ifndef BN_ACC_MUL
#define BN_ACC_MUL 4
#endif
#ifndef BN_ACC_SQR
#define BN_ACC_SQR 4
#endif
#ifndef BN_ACC_LINE
#define BN_ACC_LINE 4
#endif
#ifndef BN_ACC_CYC
#define BN_ACC_CYC 2
#endif

#ifndef BN_SMALL_CODE
#define BN_SMALL_CODE 0
#endif
#define BN_PRAGMA_(x) _Pragma(#x)
#define BN_UNROLL_N(n) BN_PRAGMA_(unroll n)
#ifndef BN_MUL_UNROLL
#define BN_MUL_UNROLL 1   // unroll factor of the six-round dense-product loop
#endif
#ifndef BN_SQR_UNROLL
#define BN_SQR_UNROLL 1
#endif
#ifndef BN_LINE_UNROLL
#define BN_LINE_UNROLL 1
#endif
#if BN_SMALL_CODE
// out-of-line modular add/sub for the hexad operations' glue code (instruction-cache footprint experiments)
BN_HD_NOINLINE Fp2 fp2_add_s(Fp2 a, Fp2 b) { return fp2_add(a, b); }
BN_HD_NOINLINE Fp2 fp2_sub_s(Fp2 a, Fp2 b) { return fp2_sub(a, b); }
#else
BN_HD Fp2 fp2_add_s(const Fp2& a, const Fp2& b) { return fp2_add(a, b); }
BN_HD Fp2 fp2_sub_s(const Fp2& a, const Fp2& b) { return fp2_sub(a, b); }
#endif
// one shared copy of the two Montgomery reductions (code footprint: the hot loops must stay inside the 32 KB I-cache)
BN_HD_NOINLINE Fp2 reduce2_wide(Wide a0, Wide a1) { return Fp2{mont_reduce<MQ, 4>(a0), mont_reduce<MQ, 4>(a1)}; }
template <class ACC>
BN_HD Fp2 reduce2(const ACC& A) {
    Wide a0, a1;
    acck_finish(A, a0, a1);
    return reduce2_wide(a0, a1);
}

BN_HD int nib(uint32_t packed, int k) { return (int)((packed >> (4 * k)) & 7u); }
BN_HD int mod6(int x) { return x >= 6 ? x - 6 : x; }

// the multiplicative identity: lane 0 holds 1
template <class Ctx>
BN_HD Fp2 hx_one(const Ctx& c) {
    return fp2_select(c.k() == 0, fp2_one(), fp2_zero());
}

// conjugation over Fq6 (w -> -w) == reference unitary_inverse, src/fields/fq12.rs:103-105
template <class Ctx>
BN_HD Fp2 hx_conj(const Ctx& c, const Fp2& a) {
    return fp2_select((c.k() & 1) != 0, fp2_neg(a), a);
}

// dense product.  reference src/fields/fq12.rs:295-307
template <class Ctx>
BN_HD_NOINLINE Fp2 hx_mul(const Ctx c, Fp2 a, Fp2 b) {
    const int k = c.k();
#if BN_F52
    {
        const Fp2 xa = c.mul_xi(a);
        c.sync();
        c.put52(0, a);
        c.put52(1, xa);
        c.put52(2, b);
        c.sync();
    }
    Acc52 acc;
    acc52_init(acc);
#if defined(__CUDA_ARCH__)
BN_UNROLL_N(BN_MUL_UNROLL)
#endif
    for (int s = 0; s < 6; s++)  // receiver k takes a_j from j = k - s (mod 6); (j, s) wraps past w^5 iff s > k
        mac52_rr(c, acc, c.ref(mod6(k + 6 - s), s > k ? 1 : 0), c.ref(s, 2));
    return acc52_reduce<6>(acc);
#else
    c.sync();
    c.put(0, a);
    c.put(1, c.mul_xi(a));
    c.put(2, b);
    c.sync();
    typename AccSel<BN_ACC_MUL>::type acc;
    acck_init(acc);
#if defined(__CUDA_ARCH__)
BN_UNROLL_N(BN_MUL_UNROLL)
#endif
    for (int s = 0; s < 6; s++) {
        // receiver k takes a_j from j = k - s (mod 6); the pair (j, s) wraps past w^5 iff j + s >= 6  <=>  s > k
        Fp2 x = c.get(mod6(k + 6 - s), s > k ? 1 : 0);
        Fp2 y = c.get(s, 2);
        mac_fp2(acc, x, y);
    }
    return reduce2(acc);
#endif
}

// square.  reference src/fields/fq12.rs:275-282.  21 distinct products in 4 lock-step rounds:
//   rounds 0-1: cross terms (the sender ships 2 a_i or 2 xi a_i)
//   round 2   : even lanes a_i^2, odd lanes their third cross term
//   round 3   : even lanes xi * a_j^2, odd lanes idle
template <class Ctx>
BN_HD_NOINLINE Fp2 hx_sqr(const Ctx c, Fp2 a) {
    const int k = c.k();
    {
        Fp2 xa = c.mul_xi(a);
        c.sync();
#if BN_F52_SQR
        c.put52(0, a);
        c.put52(1, xa);
        c.put52(2, fp2_dbl(fp2_select(k >= 4, xa, a)));
#else
        c.put(0, a);
        c.put(1, xa);
        c.put(2, fp2_dbl(fp2_select(k >= 4, xa, a)));  // doubled operand: 2 a_k on lanes 0..3, 2 xi a_k on lanes 4,5
#endif
        c.sync();
    }
#if BN_F52_SQR
    Acc52 acc;
    acc52_init(acc);
#else
    typename AccSel<BN_ACC_SQR>::type acc;
    acck_init(acc);
#endif
#if defined(__CUDA_ARCH__)
BN_UNROLL_N(BN_SQR_UNROLL)
#endif
    for (int r = 0; r < 4; r++) {
        // x sources / y sources per lane (nibble k), per round:
        //  r0: (2xi a5) a1 | (2a0) a1 | (2a0) a2 | (2a0) a3 | (2a0) a4 | (2a0) a5
        //  r1: (2xi a4) a2 | (2xi a5) a2 | (2xi a5) a3 | (2a1) a2 | (2a1) a3 | (2a1) a4
        //  r2: a0 a0 | (2xi a4) a3 | a1 a1 | (2xi a4) a5 | a2 a2 | (2a3) a2
        //  r3: xi a3 a3 | - | xi a4 a4 | - | xi a5 a5 | -
        const uint32_t xs = r == 0 ? 0x000005u : r == 1 ? 0x111554u : r == 2 ? 0x324140u : 0x050403u;
        const uint32_t ys = r == 0 ? 0x543211u : r == 1 ? 0x432322u : r == 2 ? 0x225130u : 0x050403u;
        const int xsrc = nib(xs, k);
        // slot: doubled (2) in rounds 0-1 and for sources 3,4 in round 2; plain (0) otherwise in round 2; xi (1) in round 3
        const int xslot = r < 2 ? 2 : (r == 2 ? ((xsrc == 3 || xsrc == 4) ? 2 : 0) : 1);
#if BN_F52_SQR
        mac52_rr(c, acc, c.ref(xsrc, xslot), c.ref_or_zero(!(r == 3 && (k & 1) != 0), nib(ys, k), 0));
#else
        Fp2 x = c.get(xsrc, xslot);
        Fp2 y = c.get_or_zero(!(r == 3 && (k & 1) != 0), nib(ys, k), 0);  // odd lanes idle in round 3: zero by address
        mac_fp2(acc, x, y);
#endif
    }
#if BN_F52_SQR
    return acc52_reduce<4>(acc);
#else
    return reduce2(acc);
#endif
}

// product with the sparse line l0 + l3 w^3 + l4 w^4 (reference mul_by_024, src/fields/fq12.rs:107-176).
// The coefficients are read round by round from the line source (src.coef(h, 0/1/2) = l0, l3k, l4k, where l3k / l4k are
// ALREADY the variant this lane needs:  l3k = (k < 3 ? xi*l3 : l3), l4k = (k < 4 ? xi*l4 : l4)): on the device they sit
// in the TMA ring in shared memory, so they never occupy registers next to the 512-bit accumulators.
template <class Ctx, class LineSrc>
BN_HD_NOINLINE Fp2 hx_mul_line(const Ctx c, Fp2 a, const LineSrc src, typename LineSrc::Handle h) {
    const int k = c.k();
    c.sync();
#if BN_F52_LINE
    c.put52(0, a);
    c.sync();
    Acc52 acc;
    acc52_init(acc);
#if defined(__CUDA_ARCH__)
BN_UNROLL_N(BN_LINE_UNROLL)
#endif
    for (int r = 0; r < 3; r++)  // a_k, a_{k-3}, a_{k-4}; the line coefficient is converted by the receiver
        mac52_rv(c, acc, c.ref(mod6(k + (r == 0 ? 0 : r == 1 ? 3 : 2)), 0), f52_from_fp2(src.coef(h, r)));
    return acc52_reduce<3>(acc);
#else
    c.put(0, a);
    c.sync();
    typename AccSel<BN_ACC_LINE>::type acc;
    acck_init(acc);
#if defined(__CUDA_ARCH__)
BN_UNROLL_N(BN_LINE_UNROLL)
#endif
    for (int r = 0; r < 3; r++) {
        Fp2 x = c.get(mod6(k + (r == 0 ? 0 : r == 1 ? 3 : 2)), 0);  // a_k, a_{k-3}, a_{k-4}
        Fp2 y = src.coef(h, r);
        mac_fp2(acc, x, y);
    }
    return reduce2(acc);
#endif
}

// product with an Fq6 element m0 + m1 v + m2 v^2 = m0 + m1 w^2 + m2 w^4 known to every lane.
template <class Ctx>
BN_HD_NOINLINE Fp2 hx_mul_fq6(const Ctx c, Fp2 a, Fp2 m0, Fp2 m1, Fp2 m2) {
    const int k = c.k();
    Fp2 m1k = fp2_select(k < 2, c.mul_xi(m1), m1);
    Fp2 m2k = fp2_select(k < 4, c.mul_xi(m2), m2);
    c.sync();
#if BN_F52
    c.put52(0, a);
    c.sync();
    Acc52 acc;
    acc52_init(acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 0; r < 3; r++)  // a_k, a_{k-2}, a_{k-4}
        mac52_rv(c, acc, c.ref(mod6(k + (r == 0 ? 0 : r == 1 ? 4 : 2)), 0), f52_from_fp2(fp2_select(r == 0, m0, fp2_select(r == 1, m1k, m2k))));
    return acc52_reduce<3>(acc);
#else
    c.put(0, a);
    c.sync();
    typename AccSel<BN_ACC_LINE>::type acc;
    acck_init(acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 0; r < 3; r++) {
        Fp2 x = c.get(mod6(k + (r == 0 ? 0 : r == 1 ? 4 : 2)), 0);  // a_k, a_{k-2}, a_{k-4}
        Fp2 y = fp2_select(r == 0, m0, fp2_select(r == 1, m1k, m2k));
        mac_fp2(acc, x, y);
    }
    return reduce2(acc);
#endif
}

// Frobenius x -> x^(q^p), p in {1,2,3}: conj^p on each coefficient, times xi^(k (q^p-1)/6).
// reference src/fields/fq12.rs:90-95 with the tables at fq12.rs:7-24, fq6.rs:5-40.
template <class Ctx>
BN_HD Fp2 hx_frob(const Ctx& c, const Fp2& a, int p) {
    Fp2 t = (p & 1) ? fp2_conj(a) : a;
    return fp2_mul(t, FROB_GAMMA_C[p - 1][c.k()]);
}

// Granger-Scott squaring of an element of the cyclotomic subgroup (reference src/fields/fq12.rs:178-227).
// Fq12 = Fq4[w]/(w^3 - s), Fq4 = Fq2[s]/(s^2 - xi), s = w^3: the three Fq4 coefficients are the lane pairs
// (0,3), (1,4), (2,5).  Each pair is squared with one Fq2 product per lane
//   (x + y s)^2 = [(x+y)(x + xi y) - xy - xi xy] + [2xy] s .
// Post-processing of one Fq component: with t = (pre: r - tmp - xo | lane 1: 2 xo | lanes 3,5: 2 r) the output is
// 3t -+ 2a = 2(t -+ a) + t (statement order of fq12.rs:198-221).  Everything is summed as a lazy 9-limb integer
// (negations are q - x) and reduced once: T <= 3q, Z = T + (q - a | a) <= 4q, H = 2Z + T <= 11q < 16q.
template <class Ctx>
BN_HD Fp hx_cyc_combine(const Ctx& c, const ModRegs& q, bool pre, bool lane1, const Fp& r, const Fp& tmp, const Fp& xo, const Fp& a) {
    const Fp x1 = fp_select(lane1, xo, r);
    const Fp x2 = fp_select(pre, fp_neg_lazy_r(q, tmp), x1);
    const Fp x3 = fp_select(pre, fp_neg_lazy_r(q, xo), fp_zero());
    Lazy9 t = lazy_add(x1, x2);
    lazy_acc(t, x3);
    Lazy9 z = t;
    lazy_acc(z, fp_select(pre, fp_neg_lazy_r(q, a), a));
    return c.small_reduce(lazy_dbl_add(z, t));
}
template <class Ctx>
BN_HD_NOINLINE Fp2 hx_cyc_sqr(const Ctx c, Fp2 a) {
    const int k = c.k();
    const bool pre = (k & 1) == 0;  // lanes 0,2,4 form (x+y)(x+xi y); lanes 3,5,1 form x*y
    // pair assignment: lane0,3 <- (g0,g3); lane2,5 <- (g1,g4); lane4,1 <- (g2,g5)
    const int lo = nib(0x120120u, k), hi = lo + 3;
    c.sync();
    c.put(0, a);
    c.put(1, c.mul_xi(a));
    c.sync();
    // factors (components < 2q): pre lanes (x + y)(x + xi y), the others x * y; the role picks the ADDRESS it loads
    // from (a slot or the block's zero slot), so no selects are spent on operands
    Fp2 f0 = c.get(lo, 0), f1 = c.get(hi, pre ? 1 : 0);
    {
        const Fp2 u = c.get_or_zero(pre, hi, 0), v = c.get_or_zero(pre, lo, 0);
        f0.c0 = fp_add_raw(f0.c0, u.c0);
        f0.c1 = fp_add_raw(f0.c1, u.c1);
        f1.c0 = fp_add_raw(f1.c0, v.c0);
        f1.c1 = fp_add_raw(f1.c1, v.c1);
    }
#if BN_F52_CYC
    Acc52 acc;
    acc52_init(acc);
    acc52_mac(acc, f52_from_fp2(f0), f52_from_fp2(f1));  // components < 2q, their sums < 4q < 2^256
    Fp2 r = acc52_reduce<1>(acc);
#else
    typename AccSel<BN_ACC_CYC>::type acc;
    acck_init(acc);
    mac_fp2(acc, f0, f1);
    Fp2 r = reduce2(acc);
#endif
    // partner product: lane0 <- lane3, lane2 <- lane5, lane4 <- lane1
    c.put(2, r);
    c.sync();
    Fp2 tmp = c.get(nib(0x010503u, k), 2);
    // one xi-multiplication serves both roles: xi*tmp on the "pre" lanes, xi*r on lane 1
    Fp2 xo = c.mul_xi(fp2_select(pre, tmp, r));
    // xi*r is complex-valued: component 0 of the output uses component 0 of every term, component 1 likewise
    const ModRegs q = c.mod_q();
    return Fp2{hx_cyc_combine(c, q, pre, k == 1, r.c0, tmp.c0, xo.c0, a.c0), hx_cyc_combine(c, q, pre, k == 1, r.c1, tmp.c1, xo.c1, a.c1)};
}

// f^u then conjugate (reference exp_by_neg_z, src/fields/fq12.rs:97-101, 229-246).
// VALID FOR ELEMENTS OF THE CYCLOTOMIC SUBGROUP ONLY (all three uses inside the final exponentiation are):
// there f^-1 = conj(f), so u is walked in width-4 NAF (digits 0, +-1, +-3, +-5, +-7): 62 Granger-Scott squarings and
// 13 + 3 multiplications (+ 1 squaring for a^2) instead of the reference's 62 + 27 (width 3: 17 + 1).  Gt is canonical,
// so any addition chain for the same exponent gives the same bytes.
// Register diet: the odd powers a, a^3, a^5, a^7 are only touched on the 14 non-zero digits, so they are PARKED in the
// lane's private shared-memory slots (c.park / c.unpark, slots HX_PARK_A + 0..3) and only the running value stays in
// registers.  On return slot HX_PARK_A still holds a (hx_final_exp re-reads it).
#define HX_PARK_A 0   // a, a^3, a^5, a^7 in slots 0..3
#define HX_PARK_X 4   // two slots for the caller's values that are idle during an exponentiation
#define HX_PARK_Y 5
#define HX_PARK_SLOTS 6
template <class Ctx>
BN_HD_NOINLINE Fp2 hx_exp_by_neg_z(const Ctx c, Fp2 a) {
    c.park(HX_PARK_A, a);
    {
        Fp2 a2 = hx_cyc_sqr(c, a);
        Fp2 p = a;
        for (int j = 1; j < 4; j++) {  // a^3, a^5, a^7
            p = hx_mul(c, p, a2);
            c.park(HX_PARK_A + j, p);
        }
    }
    Fp2 res = c.unpark(HX_PARK_A + BN_U_WNAF_TOP);  // leading digit
    for (int b = BN_U_WNAF_LEN - 2; b >= 0; b--) {
        res = hx_cyc_sqr(c, res);
        if ((BN_U_WNAF_NZ >> b) & 1ULL) {
            const int idx = (int)((BN_U_WNAF_M0 >> b) & 1ULL) | ((int)((BN_U_WNAF_M1 >> b) & 1ULL) << 1);
            Fp2 m = c.unpark(HX_PARK_A + idx);
            if ((BN_U_WNAF_NEG >> b) & 1ULL) m = hx_conj(c, m);
            res = hx_mul(c, m, res);
        }
    }
    return hx_conj(c, res);
}

// exp_by_neg_z exactly as the reference evaluates it (src/fields/fq12.rs:97-101, 229-246): MSB-first binary walk of u that
// skips the squarings before the first set bit, literal Granger-Scott squaring, `self * res`, then conjugation.  Unlike
// hx_exp_by_neg_z above it makes no use of f^-1 = conj(f), so it matches the reference on inputs OUTSIDE the cyclotomic
// subgroup too (the reference's test_cyclotomic_exp vector, src/fields/mod.rs:171-201, is one).
template <class Ctx>
BN_HD Fp2 hx_exp_by_neg_z_literal(const Ctx& c, const Fp2& a) {
    Fp2 res = hx_one(c);
    bool found_one = false;
    for (int b = 63; b >= 0; b--) {
        if (found_one) res = hx_cyc_sqr(c, res);
        if ((BN_U_PARAM >> b) & 1ULL) {
            found_one = true;
            res = hx_mul(c, a, res);
        }
    }
    return hx_conj(c, res);
}

// 1/f for f != 0.  reference src/fields/fq12.rs:284-292 -> fq6.rs:129-141 -> fq2.rs:125-136.
// f^-1 = conj(f) * N^-1 with N = f * conj(f) in Fq6; the small Fq6/Fq2 chain down to ONE Fq inversion is done
// redundantly by every lane (a handful of Fq2 products).  Split in two so that the Fq inversion itself can happen
// anywhere in between -- in place (hx_inv, via Ctx::inv), or for the whole batch at once in a separate kernel
// (k_miller prepares, k_fq_inv_batch inverts all norms with one warp-wide Montgomery trick, k_fexp finishes):
//   prepare:  u_i = conj(dd) t_i (i = 0..2) and the norm nn = |dd|^2 in Fq, where (t0, t1, t2) / dd = N^-1
//   finish :  f^-1 = conj(f) * ((u0, u1, u2) / nn)
struct HxInvPrep {
    Fp2 u0, u1, u2;
    Fp nn;
};
template <class Ctx>
BN_HD_NOINLINE HxInvPrep hx_inv_prepare(const Ctx c, Fp2 f) {
    Fp2 n = hx_mul(c, f, hx_conj(c, f));  // odd coefficients are zero
    c.sync();
    c.put(0, n);
    c.sync();
    Fp2 n0 = c.get(0, 0), n1 = c.get(2, 0), n2 = c.get(4, 0);
    // Fq6 inverse, reference src/fields/fq6.rs:129-141
    Fp2 t0 = fp2_sub(fp2_sqr(n0), fp2_mul(n1, fp2_mul_xi(n2)));
    Fp2 t1 = fp2_sub(fp2_mul_xi(fp2_sqr(n2)), fp2_mul(n0, n1));
    Fp2 t2 = fp2_sub(fp2_sqr(n1), fp2_mul(n0, n2));
    Fp2 dd = fp2_add(fp2_mul_xi(fp2_add(fp2_mul(n2, t1), fp2_mul(n1, t2))), fp2_mul(n0, t0));
    // 1/dd = conj(dd) / (d0^2 + d1^2)
    Wide nn = wide_zero();
    wide_mac2(nn, dd.c0, dd.c0, dd.c1, dd.c1);
    const Fp2 dc = fp2_conj(dd);
    return HxInvPrep{fp2_mul(dc, t0), fp2_mul(dc, t1), fp2_mul(dc, t2), mont_reduce<MQ, 2>(nn)};
}
template <class Ctx>
BN_HD Fp2 hx_inv_finish(const Ctx& c, const Fp2& f, const Fp2& u0, const Fp2& u1, const Fp2& u2, const Fp& ninv) {
    return hx_mul_fq6(c, hx_conj(c, f), fp2_mul_fp(u0, ninv), fp2_mul_fp(u1, ninv), fp2_mul_fp(u2, ninv));
}
template <class Ctx>
BN_HD_NOINLINE Fp2 hx_inv(const Ctx c, Fp2 f) {
    const HxInvPrep p = hx_inv_prepare(c, f);
    // the Fq inversion is delegated to the context so a kernel can batch it (Montgomery's trick over the hexads of a block)
    return hx_inv_finish(c, f, p.u0, p.u1, p.u2, c.inv(p.nn));
}

// reference final_exponentiation, src/fields/fq12.rs:41-88: same first chunk, and a last chunk that reaches the same
// exponent (q^6-1)(q^2+1) * 2u(6u^2+3u+1)(q^4-q^2+1)/r the crate's Gt values carry, regrouped
// so that at most two values are idle during each exponentiation by u (they are parked in shared memory; with the
// reference's grouping four Fq12 values = 64 registers per lane stay live across the third exponentiation).
// With s = first chunk, A = s^-u, B = A^2, D = B^3, E = D^-u, G = (E^2)^-u and K = conj(G) E conj(D), the
// reference computes   frob3(conj(s) K B) * frob2(K) * frob1(K B) * (s K E).   Frobenius is a ring homomorphism, so this
// equals   [frob3(conj(s) B) * frob1(B) * s E] * K * frob1(K) * frob2(K) * frob3(K)   : the bracket (C below) and
// E conj(D) are formed BEFORE the third exponentiation.  Gt elements are canonical, so the bytes are identical.
// `finv` = f^-1 (hx_inv, or prepared / finished around a batch-wide inversion).
template <class Ctx>
BN_HD Fp2 hx_final_exp_with_inverse(const Ctx& c, const Fp2& f, const Fp2& finv) {
    // first chunk
    Fp2 s;
    {
        Fp2 a = hx_conj(c, f);
        Fp2 cc = hx_mul(c, a, finv);
        Fp2 d = hx_frob(c, cc, 2);
        s = hx_mul(c, d, cc);
    }
    // last chunk
    c.park(HX_PARK_X, s);
    Fp2 a = hx_exp_by_neg_z(c, s);
    Fp2 b = hx_cyc_sqr(c, a);
    Fp2 d = hx_mul(c, hx_cyc_sqr(c, b), b);
    c.park(HX_PARK_Y, b);
    Fp2 e = hx_exp_by_neg_z(c, d);                                   // parked: s, b;  slot A holds d afterwards
    Fp2 ed = hx_mul(c, e, hx_conj(c, c.unpark(HX_PARK_A)));           // E conj(D)
    {
        Fp2 sv = c.unpark(HX_PARK_X), bv = c.unpark(HX_PARK_Y);
        Fp2 cst = hx_mul(c, hx_frob(c, hx_mul(c, hx_conj(c, sv), bv), 3), hx_frob(c, bv, 1));
        cst = hx_mul(c, cst, hx_mul(c, sv, e));                       // C
        c.park(HX_PARK_Y, cst);
    }
    c.park(HX_PARK_X, ed);
    Fp2 g = hx_exp_by_neg_z(c, hx_cyc_sqr(c, e));                    // parked: E conj(D), C
    Fp2 kk = hx_mul(c, hx_conj(c, g), c.unpark(HX_PARK_X));          // K
    Fp2 r = hx_mul(c, c.unpark(HX_PARK_Y), kk);
    r = hx_mul(c, r, hx_frob(c, kk, 1));
    r = hx_mul(c, r, hx_frob(c, kk, 2));
    return hx_mul(c, r, hx_frob(c, kk, 3));
}
template <class Ctx>
BN_HD Fp2 hx_final_exp(const Ctx& c, const Fp2& f) {
    return hx_final_exp_with_inverse(c, f, hx_inv(c, f));
}

// Gt::pow, reference src/fields/mod.rs:35-46 via src/lib.rs:171 (256 generic squarings, a multiplication per set bit).
// Same value by fixed 2-bit windows: two squarings and at most one multiplication by a, a^2 or a^3 per window (the
// product is computed by all hexads of the warp and selected per hexad; a zero window keeps the running value).  Valid for
// ANY Fq12 element (no use of the cyclotomic structure); the result is canonical, hence the reference's bytes.
// e = plain (de-Montgomerised) exponent, identical in all six lanes.
template <class Ctx>
BN_HD Fp2 hx_pow(const Ctx& c, const Fp2& a, const Fp& e) {
    const Fp2 a2 = hx_sqr(c, a);
    const Fp2 a3 = hx_mul(c, a2, a);
    Fp2 res = hx_one(c);
    for (int i = 127; i >= 0; i--) {
        uint32_t w = 0;
        BN_UNROLL
        for (int l = 0; l < 8; l++) w = ((i >> 4) == l) ? e.v[l] : w;
        const uint32_t d = (w >> ((i & 15) * 2)) & 3u;
        if (i != 127) {
            res = hx_sqr(c, res);
            res = hx_sqr(c, res);
        }
        const Fp2 m = fp2_select(d == 1, a, fp2_select(d == 2, a2, a3));
        const Fp2 prod = hx_mul(c, res, m);
        res = fp2_select(d != 0, prod, res);
    }
    return res;
}

// a^e for a in the CYCLOTOMIC SUBGROUP (every pairing value is): fixed 2-bit windows, two Granger-Scott squarings and
// one multiplication by {1, a, a^2, a^3} per window.  Same canonical result as Gt::pow's generic square-and-multiply
// (reference src/fields/mod.rs:35-46) at 45 % of its multiplier work; used by the fused pairing(...).pow(s) entry
// point (reference examples/joux.rs:19-21).  e = plain exponent, identical in the six lanes.
template <class Ctx>
BN_HD Fp2 hx_pow_cyc(const Ctx& c, const Fp2& a, const Fp& e) {
    const Fp2 a2 = hx_cyc_sqr(c, a);
    const Fp2 a3 = hx_mul(c, a2, a);
    const Fp2 one = hx_one(c);
    Fp2 res = one;
    for (int i = 127; i >= 0; i--) {
        uint32_t w = 0;
        BN_UNROLL
        for (int l = 0; l < 8; l++) w = ((i >> 4) == l) ? e.v[l] : w;
        const uint32_t d = (w >> ((i & 15) * 2)) & 3u;
        if (i != 127) {
            res = hx_cyc_sqr(c, res);
            res = hx_cyc_sqr(c, res);
        }
        Fp2 m = fp2_select(d == 0, one, fp2_select(d == 1, a, fp2_select(d == 2, a2, a3)));
        res = hx_mul(c, res, m);
    }
    return res;
}

// memory layout of bn::Gt (c[2][3][2][4] u64): coefficient g_k sits at Fq2 index (k&1)*3 + (k>>1)
BN_HD int gt_slot(int k) { return (k & 1) * 3 + (k >> 1); }

}  // namespace bn
