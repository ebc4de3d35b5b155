// fp2.cuh -- thread-local Fq2 = Fq[i]/(i^2+1) arithmetic (one thread owns the whole element).
//
// Replaces reference src/fields/fq2.rs:63-188 on canonical Montgomery values.  Used by the
// thread-per-element kernels (line precomputation, G1/G2 scalar multiplication) and for the small
// serial pieces of the hexad kernel.  Products are accumulated as 512-bit integers and reduced once
// per output coefficient (2 reductions per Fq2 mul/sqr instead of the reference's 3-4 Montgomery muls).
#pragma once
#include "fp.cuh"

namespace bn {

struct Fp2 {
    Fp c0, c1;
};

#if defined(__CUDACC__)
#define BN_CONST __constant__ const
#else
#define BN_CONST static const
#endif
#include "constants_tower.inc"

typedef ModQ MQ;

BN_HD Fp2 fp2_zero() { return Fp2{fp_zero(), fp_zero()}; }
BN_HD Fp fq_one() {
    Fp r;
    BN_UNROLL
    for (int i = 0; i < 8; i++) r.v[i] = FQ_ONE_f(i);
    return r;
}
BN_HD Fp2 fp2_one() { return Fp2{fq_one(), fp_zero()}; }
BN_HD Fp2 g2_coeff_b() {  // 3/xi, reference src/groups/mod.rs:392-397
    Fp2 r;
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        r.c0.v[i] = G2_B_f(i);
        r.c1.v[i] = G2_B_f(8 + i);
    }
    return r;
}
// 1/x (Montgomery in, Montgomery out) by the binary extended Euclid of reference src/arith.rs:281-327, followed by
// the same R^3 fix-up as reference src/fields/fp.rs:103-112.  Data-dependent loops: meant for ONE thread
// (block_batch_inv in kernels.cu), where it is ~3x shorter than the Fermat chain fp_inv; never call it warp-wide.
BN_HD_NOINLINE Fp fq_inv_euclid(Fp x) {
    uint32_t u[8], v[8], b[8], c[8], p[8], t[8];
    load_mod<MQ>(p);
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        u[i] = x.v[i];
        v[i] = p[i];
        b[i] = i == 0 ? 1u : 0u;
        c[i] = 0u;
    }
    auto is_one = [](const uint32_t* a) {
        uint32_t o = a[0] ^ 1u;
        BN_UNROLL
        for (int i = 1; i < 8; i++) o |= a[i];
        return o == 0;
    };
    auto shr1 = [](uint32_t* a) {
        BN_UNROLL
        for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
        a[7] >>= 1;
    };
    auto halve_mod = [&](uint32_t* a) {  // a/2 mod p for a < p
        if (a[0] & 1u) (void)addi8(a, p);  // < 2p < 2^255
        shr1(a);
    };
    auto sub_mod = [&](uint32_t* a, const uint32_t* s) {  // a = a - s mod p
        uint32_t bw = sub8(t, a, s);
        BN_UNROLL
        for (int i = 0; i < 8; i++) a[i] = t[i];
        if (bw) (void)addi8(a, p);
    };
    int guard = 0;
    while (!is_one(u) && !is_one(v) && guard++ < 1024) {
        while (!(u[0] & 1u)) {
            shr1(u);
            halve_mod(b);
        }
        while (!(v[0] & 1u)) {
            shr1(v);
            halve_mod(c);
        }
        uint32_t bw = sub8(t, u, v);  // u >= v ?
        if (!bw) {
            BN_UNROLL
            for (int i = 0; i < 8; i++) u[i] = t[i];
            sub_mod(b, c);
        } else {
            (void)sub8(t, v, u);
            BN_UNROLL
            for (int i = 0; i < 8; i++) v[i] = t[i];
            sub_mod(c, b);
        }
    }
    Fp r, r3;
    const bool use_b = is_one(u);
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        r.v[i] = use_b ? b[i] : c[i];
        r3.v[i] = FQ_R3_f(i);
    }
    return fp_mul<MQ>(r, r3);
}

BN_HD bool fp2_is_zero(const Fp2& a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
BN_HD bool fp2_eq(const Fp2& a, const Fp2& b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
BN_HD Fp2 fp2_select(bool c, const Fp2& a, const Fp2& b) {
    return Fp2{fp_select(c, a.c0, b.c0), fp_select(c, a.c1, b.c1)};
}
BN_HD Fp2 fp2_add(const Fp2& a, const Fp2& b) { return Fp2{fp_add<MQ>(a.c0, b.c0), fp_add<MQ>(a.c1, b.c1)}; }
BN_HD Fp2 fp2_sub(const Fp2& a, const Fp2& b) { return Fp2{fp_sub<MQ>(a.c0, b.c0), fp_sub<MQ>(a.c1, b.c1)}; }
BN_HD Fp2 fp2_neg(const Fp2& a) { return Fp2{fp_neg<MQ>(a.c0), fp_neg<MQ>(a.c1)}; }
BN_HD Fp2 fp2_dbl(const Fp2& a) { return Fp2{fp_dbl<MQ>(a.c0), fp_dbl<MQ>(a.c1)}; }
BN_HD Fp2 fp2_half(const Fp2& a) { return Fp2{fp_half<MQ>(a.c0), fp_half<MQ>(a.c1)}; }  // == scale(two_inv), groups/mod.rs:446-449
BN_HD Fp2 fp2_conj(const Fp2& a) { return Fp2{a.c0, fp_neg<MQ>(a.c1)}; }                // frobenius_map(odd), fq2.rs:74-83

// raw limb add without reduction (inputs canonical -> result < 2q < 2^255)
BN_HD Fp fp_add_raw(const Fp& a, const Fp& b) {
    Fp r;
    (void)add8(r.v, a.v, b.v);
    return r;
}
// 2q - a for a < 2q (lazy negation: result in (0, 2q], congruent to -a)
BN_HD Fp fp_neg_2q(const Fp& a) {
    Fp r;
    uint32_t p2[8];
    load_mod2<MQ>(p2);
    (void)sub8(r.v, p2, a.v);
    return r;
}

// (a0 + a1 i)(b0 + b1 i), schoolbook on 512-bit accumulators: 4 products, 2 reductions.
// reference src/fields/fq2.rs:139-155 (Karatsuba + 3-4 reductions there; same canonical result).
BN_HD_NOINLINE Fp2 fp2_mul(Fp2 a, Fp2 b) {
    Fp nb1 = fp_neg_lazy<MQ>(b.c1);
    return Fp2{fp_mul2<MQ>(a.c0, b.c0, a.c1, nb1),    // a0 b0 - a1 b1   (< 2 q^2)
               fp_mul2<MQ>(a.c0, b.c1, a.c1, b.c0)};  // a0 b1 + a1 b0
}
// (a0+a1)(a0-a1) + 2 a0 a1 i.   reference src/fields/fq2.rs:112-123
BN_HD_NOINLINE Fp2 fp2_sqr(Fp2 a) {
    Fp s = fp_add_raw(a.c0, a.c1);                      // < 2q
    Fp d = fp_add_raw(a.c0, fp_neg_lazy<MQ>(a.c1));     // a0 + (q - a1) in (0, 2q)
    // s d < 4 q^2 -> raw < 1.76 q; (2 a0) a1 < 2 q^2
    return Fp2{fp_mul<MQ>(s, d), fp_mul<MQ>(fp_add_raw(a.c0, a.c0), a.c1)};
}
// scale by an Fq element.   reference src/fields/fq2.rs:63-68
BN_HD Fp2 fp2_mul_fp(const Fp2& a, const Fp& k) { return Fp2{fp_mul<MQ>(a.c0, k), fp_mul<MQ>(a.c1, k)}; }

// Multiples k*q, k = 0..15, 9 limbs each (12-word rows): the quotient-estimate reduction below subtracts one row.
#define BN_KQ_STRIDE 12
BN_HD void kq_table_fill(uint32_t* tab, int k) {  // fill row k (callers split rows across threads)
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)MQ::m(i) * (uint32_t)k;
        tab[k * BN_KQ_STRIDE + i] = (uint32_t)c;
        c >>= 32;
    }
    tab[k * BN_KQ_STRIDE + 8] = (uint32_t)c;
    tab[k * BN_KQ_STRIDE + 9] = tab[k * BN_KQ_STRIDE + 10] = tab[k * BN_KQ_STRIDE + 11] = 0;
}

// v (9 limbs, < 16q) -> v mod q.  Quotient estimate from the top bits: h = floor(v / 2^251), qhat = floor(h*42/256)
// satisfies floor(v/q) - 1 <= qhat <= floor(v/q) for v < 16q (q / 2^251 = 6.0477..., 256/42 = 6.095...), so
// v - qhat*q lies in [0, 2q) and one conditional subtraction finishes.  `row(qhat, kq)` fetches the eight low limbs of
// qhat*q (KqRowPtr: plain pointer, host emulation; KqRowLds in kernels.cu: ld.shared.v4 from a 32-bit shared address).
struct KqRowPtr {
    const uint32_t* tab;
    BN_HD void operator()(uint32_t qhat, uint32_t* kq) const {
        const uint32_t* r = tab + qhat * BN_KQ_STRIDE;
        BN_UNROLL
        for (int i = 0; i < 8; i++) kq[i] = r[i];
    }
};
template <class Row>
BN_HD void fp_small_reduce9(uint32_t* v, uint32_t* out, const Row& row) {
    const uint32_t h = (v[8] << 5) | (v[7] >> 27);
    const uint32_t qhat = (h * 42u) >> 8;
    uint32_t kq[8], t[8];
    row(qhat, kq);
    (void)sub8(t, v, kq);  // the ninth limb of the difference is zero (value < 2q < 2^255)
    cond_sub_p<MQ>(t);
    BN_UNROLL
    for (int i = 0; i < 8; i++) out[i] = t[i];
}
// table-free variant (thread-per-element kernels): binary search over 8q, 4q, 2q, q
struct KqNone {};
BN_HD void fp_small_reduce9(uint32_t* v, uint32_t* out, const KqNone&) {
    BN_UNROLL
    for (int sh = 3; sh >= 0; sh--) {
        uint32_t kq[9], t[9];
        BN_UNROLL
        for (int i = 0; i < 9; i++) {
            uint32_t lo = (i < 8) ? MQ::m(i) : 0u;
            uint32_t prev = (i > 0) ? MQ::m(i - 1) : 0u;
            kq[i] = (sh == 0) ? lo : ((lo << sh) | (prev >> (32 - sh)));
        }
        uint32_t bw = sub8(t, v, kq);
        int32_t top = (int32_t)v[8] - (int32_t)kq[8] - (int32_t)bw;  // all three are tiny
        bool neg = top < 0;
        t[8] = (uint32_t)top;
        BN_UNROLL
        for (int i = 0; i < 9; i++) v[i] = neg ? v[i] : t[i];
    }
    BN_UNROLL
    for (int i = 0; i < 8; i++) out[i] = v[i];
}

// ---- lazy 9-limb integers (value < 16q): sums of a few canonical / lazily negated elements, brought back to canonical
// form by ONE quotient-estimate reduction instead of a conditional correction per addition -------------------------------
struct Lazy9 {
    uint32_t v[9];
};
BN_HD Lazy9 lazy_add(const Fp& a, const Fp& b) {
    Lazy9 r;
    r.v[8] = add8(r.v, a.v, b.v);
    return r;
}
BN_HD void lazy_acc(Lazy9& r, const Fp& a) { r.v[8] += addi8(r.v, a.v); }
// r = 2*z + t
BN_HD Lazy9 lazy_dbl_add(const Lazy9& z, const Lazy9& t) {
    Lazy9 r;
    r.v[0] = z.v[0] << 1;
    BN_UNROLL
    for (int i = 1; i < 9; i++) r.v[i] = (z.v[i] << 1) | (z.v[i - 1] >> 31);
    r.v[8] += t.v[8] + addi8(r.v, t.v);
    return r;
}
template <class Row>
BN_HD Fp lazy_reduce(Lazy9 x, const Row& row) {
    Fp r;
    fp_small_reduce9(x.v, r.v, row);
    return r;
}

#ifndef BN_XI_IMAD
#define BN_XI_IMAD 1   // run 20: k_fexp 3.243 -> 3.216 ms (-36 ALU, +14 IMAD.WIDE per call)
#endif
// multiply by xi = 9 + i:  (9x - y) + (9y + x) i.   reference src/fields/fq2.rs:70-72 (a full Fq2 mul there)
template <class Row>
BN_HD Fp2 fp2_mul_xi_tab(const Fp2& a, const Row& row) {
    Fp2 r;
    // component 0: 9*a0 + (q - a1)  in (0, 10q];  component 1: 9*a1 + a0 in [0, 10q)
    BN_UNROLL
    for (int comp = 0; comp < 2; comp++) {
        const Fp& x = comp == 0 ? a.c0 : a.c1;
        Fp addend = comp == 0 ? fp_neg_lazy<MQ>(a.c1) : a.c0;
        uint32_t v[9];
#if BN_XI_IMAD
        // v = addend + 9 x with the multiplier: two 4-instruction IMAD.WIDE chains on the (E, O) limb split of fp.cuh (the
        // integer ALU is the busiest pipe of the hexad kernels, fmaheavy the idlest: 8 IMAD.WIDE + 11 ALU instructions
        // instead of 27 ALU instructions)
        uint32_t E[9], O[9];
        BN_UNROLL
        for (int i = 0; i < 8; i++) {
            E[i] = addend.v[i];
            O[i] = 0;
        }
        E[8] = 0;
        O[8] = 0;
        mad_row4(E, x.v[0], x.v[2], x.v[4], x.v[6], 9u);
        mad_row4(O, x.v[1], x.v[3], x.v[5], x.v[7], 9u);
        v[0] = E[0];
        (void)add8(v + 1, E + 1, O);  // v[1..8] = E[1..8] + O[0..7]; O[8] = 0 (9 x_j < 2^36) and the total is below 2^288
#else
        // v = x << 3
        v[0] = x.v[0] << 3;
        BN_UNROLL
        for (int i = 1; i < 8; i++) v[i] = (x.v[i] << 3) | (x.v[i - 1] >> 29);
        v[8] = x.v[7] >> 29;
        uint32_t c = addi8(v, x.v);
        v[8] += c;
        c = addi8(v, addend.v);
        v[8] += c;
#endif
        fp_small_reduce9(v, comp == 0 ? r.c0.v : r.c1.v, row);
    }
    return r;
}
BN_HD_NOINLINE Fp2 fp2_mul_xi(Fp2 a) { return fp2_mul_xi_tab(a, KqNone()); }
// hexad kernels: reduction rows come from a k*q table (shared memory on the device)
template <class Row>
BN_HD_NOINLINE Fp2 fp2_mul_xi_r(Fp2 a, Row row) { return fp2_mul_xi_tab(a, row); }
BN_HD Fp2 fp2_mul_xi_t(const Fp2& a, const uint32_t* tab) { return fp2_mul_xi_r(a, KqRowPtr{tab}); }

// 1/a.   reference src/fields/fq2.rs:125-136.  a != 0.
BN_HD Fp2 fp2_inv(const Fp2& a) {
    Wide n = wide_zero();
    wide_mac2(n, a.c0, a.c0, a.c1, a.c1);
    Fp t = fp_inv<MQ>(mont_reduce<MQ, 2>(n));
    return Fp2{fp_mul<MQ>(a.c0, t), fp_neg<MQ>(fp_mul<MQ>(a.c1, t))};
}

}  // namespace bn
