// curve.cuh -- Jacobian G1/G2 arithmetic and the optimal-ate line schedule, thread-per-element.
//
// Group law: the formulas and the control flow of reference src/groups/mod.rs:228-312 are kept exactly
// (same doubling / addition formulas, same early returns, same MSB-first chain that skips leading zeros),
// because the crate's G1/G2 values are un-normalised Jacobian triples: only the same chain of canonical
// field operations reproduces the same (x, y, z) limbs.  The field arithmetic underneath is ours (fp.cuh).
//
// Line schedule: the 102 line evaluations of the ate loop (64 doublings + 36 additions + 2 Frobenius
// additions; reference src/groups/mod.rs:557-634) are produced by ONE thread per pairing and streamed to
// HBM in the order the Miller kernel consumes them, already multiplied by P's coordinates and by xi where
// a lane of the Miller kernel needs the wrapped coefficient.  (The reference keeps them in a heap Vec.)
#pragma once
#include "duo.cuh"
#include "fp2.cuh"

namespace bn {

struct FqOps {
    typedef Fp T;
    static BN_HD T add(const T& a, const T& b) { return fp_add<MQ>(a, b); }
    static BN_HD T sub(const T& a, const T& b) { return fp_sub<MQ>(a, b); }
    static BN_HD T neg(const T& a) { return fp_neg<MQ>(a); }
    static BN_HD T mul(const T& a, const T& b) { return fp_mul_ni<MQ>(a, b); }
    static BN_HD T sqr(const T& a) { return fp_sqr_ni<MQ>(a); }
    static BN_HD bool is_zero(const T& a) { return fp_is_zero(a); }
    static BN_HD bool eq(const T& a, const T& b) { return fp_eq(a, b); }
    static BN_HD T zero() { return fp_zero(); }
    static BN_HD T one() { return fq_one(); }
};
struct Fq2Ops {
    typedef Fp2 T;
    static BN_HD T add(const T& a, const T& b) { return fp2_add(a, b); }
    static BN_HD T sub(const T& a, const T& b) { return fp2_sub(a, b); }
    static BN_HD T neg(const T& a) { return fp2_neg(a); }
    static BN_HD T mul(const T& a, const T& b) { return fp2_mul(a, b); }
    static BN_HD T sqr(const T& a) { return fp2_sqr(a); }
    static BN_HD bool is_zero(const T& a) { return fp2_is_zero(a); }
    static BN_HD bool eq(const T& a, const T& b) { return fp2_eq(a, b); }
    static BN_HD T zero() { return fp2_zero(); }
    static BN_HD T one() { return fp2_one(); }
};

template <class F>
struct Jac {
    typename F::T x, y, z;
};

// Doubling.  The crate's G1 / G2 values are un-normalised Jacobian triples, so limb equality with the reference needs the
// SAME sequence of canonical field operations as reference src/groups/mod.rs:228-247 (dbl-2009-l shape: X^2, Y^2, Y^4,
// S = 2((X + Y^2)^2 - X^2 - Y^4), M = 3 X^2, X' = M^2 - 2S, Y' = M (S - X') - 8 Y^4, Z' = 2 Y Z); only the names are ours.
template <class F>
BN_HD_NOINLINE Jac<F> jac_double(const Jac<F> p) {
    typedef typename F::T T;
    const T xx = F::sqr(p.x);
    const T yy = F::sqr(p.y);
    const T yyyy = F::sqr(yy);
    T s = F::sub(F::sub(F::sqr(F::add(p.x, yy)), xx), yyyy);
    s = F::add(s, s);
    const T m = F::add(F::add(xx, xx), xx);
    const T mm = F::sqr(m);
    const T xr = F::sub(mm, F::add(s, s));
    T y4x8 = F::add(yyyy, yyyy);
    y4x8 = F::add(y4x8, y4x8);
    y4x8 = F::add(y4x8, y4x8);
    const T yz = F::mul(p.y, p.z);
    Jac<F> r;
    r.x = xr;
    r.y = F::sub(F::mul(m, F::sub(s, xr)), y4x8);
    r.z = F::add(yz, yz);
    return r;
}

// Addition, same contract: the operation sequence of reference src/groups/mod.rs:272-312 (add-2007-bl shape), including
// its early exits: an operand at infinity returns the other one, equal operands fall back to the doubling, and P + (-P)
// runs through the general formula (giving a triple with z = 0 that is not the canonical (0, 1, 0)).
template <class F>
BN_HD_NOINLINE Jac<F> jac_add(const Jac<F> p, const Jac<F> o) {
    typedef typename F::T T;
    if (F::is_zero(p.z)) return o;
    if (F::is_zero(o.z)) return p;
    const T zz1 = F::sqr(p.z);
    const T zz2 = F::sqr(o.z);
    const T ua = F::mul(p.x, zz2);
    const T ub = F::mul(o.x, zz1);
    const T zzz1 = F::mul(p.z, zz1);
    const T zzz2 = F::mul(o.z, zz2);
    const T sa = F::mul(p.y, zzz2);
    const T sb = F::mul(o.y, zzz1);
    if (F::eq(ua, ub) && F::eq(sa, sb)) return jac_double<F>(p);
    const T dx = F::sub(ub, ua);
    const T dy = F::sub(sb, sa);
    const T fourdx2 = F::sqr(F::add(dx, dx));
    const T cube = F::mul(dx, fourdx2);
    const T slope2 = F::add(dy, dy);
    const T base = F::mul(ua, fourdx2);
    const T sac = F::mul(sa, cube);
    const T xr = F::sub(F::sub(F::sqr(slope2), cube), F::add(base, base));
    Jac<F> out;
    out.x = xr;
    out.y = F::sub(F::mul(slope2, F::sub(base, xr)), F::add(sac, sac));
    out.z = F::mul(F::sub(F::sub(F::sqr(F::add(p.z, o.z)), zz1), zz2), dx);
    return out;
}

// reference src/groups/mod.rs:314-327: zero is returned unchanged, otherwise (x, -y, z)
template <class F>
BN_HD Jac<F> jac_neg(const Jac<F>& p) {
    Jac<F> r = p;
    if (!F::is_zero(p.z)) r.y = F::neg(p.y);
    return r;
}

// projective equality, reference src/groups/mod.rs:83-109
template <class F>
BN_HD bool jac_eq(const Jac<F>& p, const Jac<F>& o) {
    typedef typename F::T T;
    if (F::is_zero(p.z)) return F::is_zero(o.z);
    if (F::is_zero(o.z)) return false;
    T zz1 = F::sqr(p.z), zz2 = F::sqr(o.z);
    if (!F::eq(F::mul(p.x, zz2), F::mul(o.x, zz1))) return false;
    return F::eq(F::mul(p.y, F::mul(o.z, zz2)), F::mul(o.y, F::mul(p.z, zz1)));
}

// `G * Fr`: reference src/groups/mod.rs:250-270.  fr is the Montgomery image of the scalar (as stored in bn::Fr).
template <class F>
BN_HD Jac<F> jac_mul(const Jac<F>& p, const Fp& fr) {
    Fp k = fp_from_mont<ModR>(fr);  // U256::from(Fr), reference src/fields/fp.rs:15-22
    Jac<F> res;
    res.x = F::zero();
    res.y = F::one();
    res.z = F::zero();  // G::zero(), reference src/groups/mod.rs:208-214
    bool found_one = false;
    for (int i = 255; i >= 0; i--) {
        if (found_one) res = jac_double<F>(res);
        uint32_t w = 0;
        BN_UNROLL
        for (int l = 0; l < 8; l++) w = ((i >> 5) == l) ? k.v[l] : w;
        if ((w >> (i & 31)) & 1u) {
            found_one = true;
            res = jac_add<F>(res, p);
        }
    }
    return res;
}

// ------------------------------------------------------------------------------------------------
// Optimal-ate line schedule
// ------------------------------------------------------------------------------------------------
struct G2Proj {
    Fp2 x, y, z;
};
// One line of the Miller loop as the hexad kernel consumes it.  As an Fq12 element (w^6 = xi):
//   l0 + l3 w^3 + l4 w^4,   l0 = ell_0, l3 = ell_vw * P.y, l4 = ell_vv * P.x
// (reference mul_by_024 slots, src/fields/fq12.rs:107-120, and the scaling at src/groups/mod.rs:502),
// plus xi*l3 and xi*l4 for the lanes whose product index wraps past w^5.
struct Line {
    Fp2 l0, l3, xl3, l4, xl4;
};
#define BN_LINE_WORDS 80
// Miller-loop schedule: signed-digit (NAF) walk of 6u+2 by default -- 65 doublings + 21 additions/subtractions + the 2
// Frobenius additions = 88 lines instead of the reference's binary walk (64 + 36 + 2 = 102, src/groups/mod.rs:560-582).
// Different addition chains change the Miller value only by factors from proper subfields, which the final
// exponentiation kills, so Gt is identical.  BN_ATE_NAF=0 reproduces the reference's chain exactly (the host-emulator
// tests use it to compare every line and the unreduced Miller value with the reference's known answers).
#ifndef BN_ATE_NAF
#define BN_ATE_NAF 1
#endif
#if BN_ATE_NAF
#define BN_NUM_LINES BN_NUM_LINES_NAF
#else
#define BN_NUM_LINES BN_NUM_LINES_BIN
#endif

// Fq2 multiplication policies for to_affine and the one-thread line schedule; the lane-pair schedule (duo.cuh) has its
// own step functions below (line_double_duo, line_add_duo) and uses DuoX for to_affine only.
struct SoloX {
    BN_HD Fp2 mul(const Fp2& a, const Fp2& b) const { return fp2_mul(a, b); }
    BN_HD Fp2 sqr(const Fp2& a) const { return fp2_sqr(a); }
    BN_HD Fp2 mul_fp(const Fp2& a, const Fp& k) const { return fp2_mul_fp(a, k); }
    BN_HD Fp2 mul_xi(const Fp2& a) const { return fp2_mul_xi(a); }
};
template <class D>
struct DuoX {  // whole Fq2 values in, whole values out (both lanes), computed by the pair
    D d;
    BN_HD Fp2 mul(const Fp2& a, const Fp2& b) const { return duo_whole(d, duo_mul(d, duo_own(d, a), duo_own(d, b))); }
    BN_HD Fp2 sqr(const Fp2& a) const { return duo_whole(d, duo_sqr(d, duo_own(d, a))); }
};

template <class X>
BN_HD Line make_line(const X& X_, const Fp2& ell_0, const Fp2& ell_vw, const Fp2& ell_vv, const Fp& px, const Fp& py) {
    Line L;
    L.l0 = ell_0;
    L.l3 = X_.mul_fp(ell_vw, py);
    L.l4 = X_.mul_fp(ell_vv, px);
    L.xl3 = X_.mul_xi(L.l3);
    L.xl4 = X_.mul_xi(L.l4);
    return L;
}

// reference src/groups/mod.rs:612-634
template <class X>
BN_HD_NOINLINE Line line_double(const X& X_, G2Proj& r, const Fp& px, const Fp& py) {
    Fp2 a = fp2_half(X_.mul(r.x, r.y));
    Fp2 b = X_.sqr(r.y);
    Fp2 c = X_.sqr(r.z);
    Fp2 d = fp2_add(fp2_add(c, c), c);
    Fp2 e = X_.mul(g2_coeff_b(), d);
    Fp2 f = fp2_add(fp2_add(e, e), e);
    Fp2 g = fp2_half(fp2_add(b, f));
    Fp2 h = fp2_sub(X_.sqr(fp2_add(r.y, r.z)), fp2_add(b, c));
    Fp2 i = fp2_sub(e, b);
    Fp2 j = X_.sqr(r.x);
    Fp2 e_sq = X_.sqr(e);
    r.x = X_.mul(a, fp2_sub(b, f));
    r.y = fp2_sub(X_.sqr(g), fp2_add(fp2_add(e_sq, e_sq), e_sq));
    r.z = X_.mul(b, h);
    return make_line(X_, X_.mul_xi(i), fp2_neg(h), fp2_add(fp2_add(j, j), j), px, py);
}

// reference src/groups/mod.rs:592-610
template <class X>
BN_HD_NOINLINE Line line_add(const X& X_, G2Proj& r, const Fp2& bx, const Fp2& by, const Fp& px, const Fp& py) {
    Fp2 d = fp2_sub(r.x, X_.mul(r.z, bx));
    Fp2 e = fp2_sub(r.y, X_.mul(r.z, by));
    Fp2 f = X_.sqr(d);
    Fp2 g = X_.sqr(e);
    Fp2 h = X_.mul(d, f);
    Fp2 i = X_.mul(r.x, f);
    Fp2 j = fp2_sub(fp2_add(X_.mul(r.z, g), h), fp2_add(i, i));
    r.x = X_.mul(d, j);
    r.y = fp2_sub(X_.mul(e, fp2_sub(i, j)), X_.mul(h, r.y));
    r.z = X_.mul(r.z, h);
    Fp2 ell_0 = X_.mul_xi(fp2_sub(X_.mul(e, bx), X_.mul(d, by)));
    return make_line(X_, ell_0, d, fp2_neg(e), px, py);
}

// The same two steps on a lane pair (duo.cuh).  Every value is the lane's OWN component (component h) of the Fq2 value
// of the same name above and is computed by the same formula, so lines and running point are bit-identical.
struct G2ProjH {
    Fp x, y, z;
};
struct LineH {
    Fp l0, l3, xl3, l4, xl4;
};
template <class D>
BN_HD LineH make_line_duo(const D& d, const Fp& ell_0_pre, const Fp& ell_vw, const Fp& ell_vv, const Fp& px, const Fp& py) {
    LineH L;
    L.l3 = duo_mul_fp(ell_vw, py);
    L.l4 = duo_mul_fp(ell_vv, px);
    L.xl3 = duo_mul_xi(d, L.l3);
    L.xl4 = duo_mul_xi(d, L.l4);
    L.l0 = duo_mul_xi(d, ell_0_pre);
    return L;
}
// reference src/groups/mod.rs:612-634
template <class D>
BN_HD_NOINLINE LineH line_double_duo(const D d, G2ProjH& r, Fp px, Fp py) {
    Fp b = duo_sqr(d, r.y);
    Fp c = duo_sqr(d, r.z);
    Fp t = duo_sqr(d, fp_add<MQ>(r.y, r.z));
    Fp j = duo_sqr(d, r.x);
    Fp a = fp_half<MQ>(duo_mul(d, r.x, r.y));
    Fp e = duo_mul(d, duo_own(d, g2_coeff_b()), fp_add<MQ>(fp_add<MQ>(c, c), c));
    Fp f = fp_add<MQ>(fp_add<MQ>(e, e), e);
    Fp g = fp_half<MQ>(fp_add<MQ>(b, f));
    Fp h = fp_sub<MQ>(t, fp_add<MQ>(b, c));
    Fp i = fp_sub<MQ>(e, b);
    Fp e_sq = duo_sqr(d, e);
    Fp g_sq = duo_sqr(d, g);
    r.x = duo_mul(d, a, fp_sub<MQ>(b, f));
    r.z = duo_mul(d, b, h);
    r.y = fp_sub<MQ>(g_sq, fp_add<MQ>(fp_add<MQ>(e_sq, e_sq), e_sq));
    return make_line_duo(d, i, fp_neg<MQ>(h), fp_add<MQ>(fp_add<MQ>(j, j), j), px, py);
}
// reference src/groups/mod.rs:592-610
template <class D>
BN_HD_NOINLINE LineH line_add_duo(const D d, G2ProjH& r, Fp bx, Fp by, Fp px, Fp py) {
    Fp dd = fp_sub<MQ>(r.x, duo_mul(d, r.z, bx));
    Fp e = fp_sub<MQ>(r.y, duo_mul(d, r.z, by));
    Fp f = duo_sqr(d, dd);
    Fp g = duo_sqr(d, e);
    Fp h = duo_mul(d, dd, f);
    Fp i = duo_mul(d, r.x, f);
    Fp j = fp_sub<MQ>(fp_add<MQ>(duo_mul(d, r.z, g), h), fp_add<MQ>(i, i));
    r.x = duo_mul(d, dd, j);
    r.y = fp_sub<MQ>(duo_mul(d, e, fp_sub<MQ>(i, j)), duo_mul(d, h, r.y));
    r.z = duo_mul(d, r.z, h);
    Fp ell_0_pre = fp_sub<MQ>(duo_mul(d, e, bx), duo_mul(d, dd, by));
    return make_line_duo(d, ell_0_pre, dd, fp_neg<MQ>(e), px, py);
}
// ate_lines (below) on a lane pair.  sink(index, LineH).
template <class D, class Sink>
BN_HD void ate_lines_duo(const D& d, const Fp& px, const Fp& py, const Fp2& qx2, const Fp2& qy2, Sink& sink) {
    const bool hi = d.h() != 0;
    const Fp qx = duo_own(d, qx2), qy = duo_own(d, qy2);
    G2ProjH r;
    r.x = qx;
    r.y = qy;
    r.z = hi ? fp_zero() : fq_one();
    int n = 0;
#if BN_ATE_NAF
    const Fp nqy = fp_neg<MQ>(qy);
    for (int b = BN_ATE_NAF_DIGITS - 1; b >= 0; b--) {
        sink(n++, line_double_duo(d, r, px, py));
        if (b < 64 && ((BN_ATE_NAF_NZ >> b) & 1ULL))
            sink(n++, line_add_duo(d, r, qx, ((BN_ATE_NAF_NEG >> b) & 1ULL) ? nqy : qy, px, py));
    }
#else
    for (int b = BN_ATE_NBITS - 1; b >= 0; b--) {
        sink(n++, line_double_duo(d, r, px, py));
        if ((BN_ATE_BITS >> b) & 1ULL) sink(n++, line_add_duo(d, r, qx, qy, px, py));
    }
#endif
    // twisted Frobenius (src/groups/mod.rs:550-555); conj negates component 1
    const Fp gx = duo_own(d, FROB_GAMMA_C[0][2]), gy = duo_own(d, FROB_GAMMA_C[0][3]);
    Fp q1x = duo_mul(d, gx, hi ? fp_neg<MQ>(qx) : qx), q1y = duo_mul(d, gy, hi ? fp_neg<MQ>(qy) : qy);
    Fp q2x = duo_mul(d, gx, hi ? fp_neg<MQ>(q1x) : q1x), q2y = fp_neg<MQ>(duo_mul(d, gy, hi ? fp_neg<MQ>(q1y) : q1y));
    sink(n++, line_add_duo(d, r, q1x, q1y, px, py));
    sink(n++, line_add_duo(d, r, q2x, q2y, px, py));
}

// twisted Frobenius: reference src/groups/mod.rs:550-555
template <class X>
BN_HD void g2_mul_by_q(const X& X_, Fp2& x, Fp2& y) {
    x = X_.mul(FROB_GAMMA_C[0][2], fp2_conj(x));
    y = X_.mul(FROB_GAMMA_C[0][3], fp2_conj(y));
}

// Affine coordinates of a (G1, G2) pair with ONE field inversion for both points
// (reference to_affine, src/groups/mod.rs:113-130, inverts P.z and Q.z separately; the values are the same).
// Returns false when either point is the point at infinity (pairing = one, src/groups/mod.rs:765-766).
// `inv` is a callable Fp -> Fp (plain Fermat inversion, or a block-wide batched inversion in the kernels).
template <class X, class Inv>
BN_HD bool pair_to_affine(const X& X_, const Inv& inv, const Jac<FqOps>& P, const Jac<Fq2Ops>& Q, Fp& px, Fp& py, Fp2& qx,
                          Fp2& qy) {
    bool inf = fp_is_zero(P.z) || fp2_is_zero(Q.z);
    Wide n = wide_zero();
    wide_mac2(n, Q.z.c0, Q.z.c0, Q.z.c1, Q.z.c1);
    Fp nq = mont_reduce<MQ, 2>(n);         // |Q.z|^2 in Fq
    Fp t = fp_mul<MQ>(P.z, nq);
    Fp tinv = inv(fp_select(inf, fq_one(), t));
    Fp pzinv = fp_mul<MQ>(tinv, nq);
    Fp nqinv = fp_mul<MQ>(tinv, P.z);
    Fp2 qzinv = Fp2{fp_mul<MQ>(Q.z.c0, nqinv), fp_neg<MQ>(fp_mul<MQ>(Q.z.c1, nqinv))};
    Fp pz2 = fp_mul<MQ>(pzinv, pzinv);
    px = fp_mul<MQ>(P.x, pz2);
    py = fp_mul<MQ>(P.y, fp_mul<MQ>(pz2, pzinv));
    Fp2 qz2 = X_.sqr(qzinv);
    qx = X_.mul(Q.x, qz2);
    qy = X_.mul(Q.y, X_.mul(qz2, qzinv));
    return !inf;
}

struct FermatInv {
    BN_HD Fp operator()(const Fp& x) const { return fp_inv<MQ>(x); }
};

// Emit the 102 lines for affine (P, Q) in Miller-loop order.  sink(index, line).
// reference precompute, src/groups/mod.rs:557-588
template <class X, class Sink>
BN_HD void ate_lines(const X& X_, const Fp& px, const Fp& py, const Fp2& qx, const Fp2& qy, Sink& sink) {
    G2Proj r;
    r.x = qx;
    r.y = qy;
    r.z = fp2_one();
    int n = 0;
#if BN_ATE_NAF
    const Fp2 nqy = fp2_neg(qy);
    for (int b = BN_ATE_NAF_DIGITS - 1; b >= 0; b--) {
        sink(n++, line_double(X_, r, px, py));
        if (b < 64 && ((BN_ATE_NAF_NZ >> b) & 1ULL))
            sink(n++, line_add(X_, r, qx, ((BN_ATE_NAF_NEG >> b) & 1ULL) ? nqy : qy, px, py));
    }
#else
    for (int b = BN_ATE_NBITS - 1; b >= 0; b--) {
        sink(n++, line_double(X_, r, px, py));
        if ((BN_ATE_BITS >> b) & 1ULL) sink(n++, line_add(X_, r, qx, qy, px, py));
    }
#endif
    Fp2 q1x = qx, q1y = qy;
    g2_mul_by_q(X_, q1x, q1y);
    Fp2 q2x = q1x, q2y = q1y;
    g2_mul_by_q(X_, q2x, q2y);
    q2y = fp2_neg(q2y);
    sink(n++, line_add(X_, r, q1x, q1y, px, py));
    sink(n++, line_add(X_, r, q2x, q2y, px, py));
}

}  // namespace bn
