// wire.cuh -- the crate's wire format for field and group elements, per element (host/device).
//
// Row f-3 of SURVEY.md section 8: the steps either side of the pairing in a verifier.  Reference:
//   Fq / Fr  : 32-byte big-endian canonical integer          src/arith.rs:128-159, src/fields/fp.rs:24-36
//   Fq2      : 64-byte big-endian integer c1 * q + c0         src/fields/fq2.rs:31-53, src/arith.rs:21-44, 65-88
//   G1 / G2  : 0x00 (infinity) | 0x04 || x || y  (affine)     src/groups/mod.rs:143-205
// Values inside the library are Montgomery form; the wire holds plain integers.
// Decoding follows the reference's checks in its order: leading byte, "integer is not less than modulus"
// (for Fq2: the 512-bit integer must be < q^2, i.e. c1 < q after U512::divrem), on-curve, and for G2 the order-r
// subgroup test p * (-1) + p == 0 (src/groups/mod.rs:178-205, check_order :399).
#pragma once
#include "curve.cuh"

namespace bn {

enum WireStatus : uint8_t {
    WIRE_OK = 0,
    WIRE_BAD_TAG = 1,        // "invalid leading byte for uncompressed group element"
    WIRE_NOT_REDUCED = 2,    // "integer is not less than modulus"
    WIRE_NOT_ON_CURVE = 3,   // "point is not on the curve"
    WIRE_NOT_IN_SUBGROUP = 4 // "point is not in the subgroup"
};

// 32 big-endian bytes <-> 8 little-endian 32-bit limbs (plain integer)
BN_HD void limbs_from_be32(const uint8_t* b, uint32_t* v) {
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        const uint8_t* p = b + 4 * (7 - i);
        v[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
}
BN_HD void limbs_to_be32(const uint32_t* v, uint8_t* b) {
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        uint8_t* p = b + 4 * (7 - i);
        p[0] = (uint8_t)(v[i] >> 24);
        p[1] = (uint8_t)(v[i] >> 16);
        p[2] = (uint8_t)(v[i] >> 8);
        p[3] = (uint8_t)v[i];
    }
}

template <class M>
BN_HD Fp mont_r2();
template <>
BN_HD Fp mont_r2<ModQ>() {
    Fp r;
    BN_UNROLL
    for (int i = 0; i < 8; i++) r.v[i] = FQ_R2_f(i);
    return r;
}
template <>
BN_HD Fp mont_r2<ModR>() {
    Fp r;
    BN_UNROLL
    for (int i = 0; i < 8; i++) r.v[i] = FR_R2_f(i);
    return r;
}

// plain canonical integer -> 32 bytes  (Montgomery in).  reference Fq/Fr encode: src/fields/fp.rs:24-29
template <class M>
BN_HD void fp_encode(const Fp& a, uint8_t* b) {
    Fp p = fp_from_mont<M>(a);
    limbs_to_be32(p.v, b);
}
// 32 bytes -> Montgomery; false when the integer is >= the modulus.  reference src/fields/fp.rs:31-36, 62-70
template <class M>
BN_HD bool fp_decode(const uint8_t* b, Fp& out) {
    Fp x;
    limbs_from_be32(b, x.v);
    uint32_t t[8], p[8];
    load_mod<M>(p);
    const bool ok = sub8(t, x.v, p) != 0;  // borrow <=> x < p
    out = fp_mul<M>(x, mont_r2<M>());       // x * R^2 / R = x * R   (x < 2^256 keeps the product below p * 2^256... see below)
    return ok;
}
// Note on fp_decode for x >= p: the product x * R^2 is < 2^256 * p, the reduction stays in range and the result is the
// Montgomery form of x mod p; the caller discards it (status WIRE_NOT_REDUCED), exactly like the reference's Err.

// Fq2 -> 64 bytes: the 512-bit integer c1 * q + c0.  reference src/fields/fq2.rs:31-40
BN_HD void fq2_encode(const Fp2& a, uint8_t* b) {
    Fp c0 = fp_from_mont<MQ>(a.c0), c1 = fp_from_mont<MQ>(a.c1), qm;
    load_mod<MQ>(qm.v);
    Wide T = wide_zero();
    BN_UNROLL
    for (int i = 0; i < 8; i++) T.w[i] = c0.v[i];
    wide_mac1(T, c1, qm);  // c1 * q + c0 < q^2 < 2^512
    limbs_to_be32(T.w + 8, b);
    limbs_to_be32(T.w, b + 32);
}
// 64 bytes -> Fq2 (Montgomery); false when the integer is >= q^2 (the reference's divrem yields c1 >= q).
// reference src/fields/fq2.rs:42-53.  N = c1 * q + c0:  c0 = N mod q through two Montgomery steps
// (N R^-1, then * R^3 R^-1 = N R: already the Montgomery form of c0), c1 = (N - c0) / q exactly, computed as
// (N - c0) * q^-1 mod 2^256 (valid because c1 < q < 2^256 once N < q^2 has been checked).
BN_HD bool fq2_decode(const uint8_t* b, Fp2& out) {
    Wide N;
    limbs_from_be32(b, N.w + 8);
    limbs_from_be32(b + 32, N.w);
    // N < q^2 ?
    uint32_t qq[16], t[8];
    BN_UNROLL
    for (int i = 0; i < 16; i++) qq[i] = Q_SQUARED_f(i);
    uint32_t bw = sub8(t, N.w, qq);
    uint32_t hi[8];
    BN_UNROLL
    for (int i = 0; i < 8; i++) hi[i] = N.w[8 + i];
    bw = subi8b(hi, qq + 8, bw);
    const bool ok = bw != 0;
    // c0 (plain) = mont(mont_reduce(N), R^2);  c0 (Montgomery) = mont(mont_reduce(N), R^3)
    Wide Nc = N;
    if (!ok) {  // keep the reduction's precondition (T < q * 2^256) for garbage input; the result is discarded
        BN_UNROLL
        for (int i = 8; i < 16; i++) Nc.w[i] = 0;
    }
    Fp nr = mont_reduce<MQ, 2>(Nc);  // N / R mod q
    Fp r3;
    BN_UNROLL
    for (int i = 0; i < 8; i++) r3.v[i] = FQ_R3_f(i);
    out.c0 = fp_mul<MQ>(nr, r3);
    Fp c0 = fp_from_mont<MQ>(out.c0);
    // d = low 256 bits of N - c0;  c1 = d * q^-1 mod 2^256
    uint32_t d[8];
    (void)sub8(d, Nc.w, c0.v);
    Fp dd, qi;
    BN_UNROLL
    for (int i = 0; i < 8; i++) {
        dd.v[i] = d[i];
        qi.v[i] = Q_INV256_f(i);
    }
    Wide P;
    wide_mul(P, dd, qi);
    Fp c1;
    BN_UNROLL
    for (int i = 0; i < 8; i++) c1.v[i] = P.w[i];
    if (!ok) c1 = fp_zero();
    out.c1 = fp_mul<MQ>(c1, mont_r2<MQ>());
    return ok;
}

// on-curve test of an affine point given in Montgomery form.  reference AffineG::new, src/groups/mod.rs:425-433
BN_HD bool g1_on_curve(const Fp& x, const Fp& y) {
    Fp b;
    BN_UNROLL
    for (int l = 0; l < 8; l++) b.v[l] = G1_B_f(l);
    return fp_eq(fp_mul<MQ>(y, y), fp_add<MQ>(fp_mul<MQ>(fp_mul<MQ>(x, x), x), b));
}
BN_HD bool g2_on_curve(const Fp2& x, const Fp2& y) {
    return fp2_eq(fp2_sqr(y), fp2_add(fp2_mul(fp2_sqr(x), x), g2_coeff_b()));
}
// p * (-1) + p == zero  (reference src/groups/mod.rs:193-199)
BN_HD bool g2_in_subgroup(const Jac<Fq2Ops>& P) {
    Fp m1;
    BN_UNROLL
    for (int l = 0; l < 8; l++) m1.v[l] = FR_MINUS_ONE_f(l);
    Jac<Fq2Ops> r = jac_add<Fq2Ops>(jac_mul<Fq2Ops>(P, m1), P);
    return fp2_is_zero(r.z);
}

// G1 record: 65 bytes.  Returns the status; *out is the Jacobian point (x, y, one) or zero() = (0, 1, 0).
BN_HD uint8_t g1_decode(const uint8_t* rec, Jac<FqOps>& out) {
    out.x = fp_zero();
    out.y = fq_one();
    out.z = fp_zero();
    if (rec[0] == 0) return WIRE_OK;
    if (rec[0] != 4) return WIRE_BAD_TAG;
    Fp x, y;
    const bool okx = fp_decode<MQ>(rec + 1, x);
    const bool oky = fp_decode<MQ>(rec + 33, y);
    if (!okx || !oky) return WIRE_NOT_REDUCED;
    if (!g1_on_curve(x, y)) return WIRE_NOT_ON_CURVE;
    out.x = x;
    out.y = y;
    out.z = fq_one();
    return WIRE_OK;
}
// G2 record: 129 bytes.
BN_HD uint8_t g2_decode(const uint8_t* rec, Jac<Fq2Ops>& out) {
    out.x = fp2_zero();
    out.y = fp2_one();
    out.z = fp2_zero();
    if (rec[0] == 0) return WIRE_OK;
    if (rec[0] != 4) return WIRE_BAD_TAG;
    Fp2 x, y;
    const bool okx = fq2_decode(rec + 1, x);
    const bool oky = fq2_decode(rec + 65, y);
    if (!okx || !oky) return WIRE_NOT_REDUCED;
    if (!g2_on_curve(x, y)) return WIRE_NOT_ON_CURVE;
    Jac<Fq2Ops> P{x, y, fp2_one()};
    if (!g2_in_subgroup(P)) return WIRE_NOT_IN_SUBGROUP;
    out = P;
    return WIRE_OK;
}
// affine (already normalised: z == one or z == 0) -> record; infinity writes 0x00 followed by zero padding
BN_HD void g1_encode_affine(const Fp& x, const Fp& y, bool infinity, uint8_t* rec) {
    if (infinity) {
        for (int i = 0; i < 65; i++) rec[i] = 0;
        return;
    }
    rec[0] = 4;
    fp_encode<MQ>(x, rec + 1);
    fp_encode<MQ>(y, rec + 33);
}
BN_HD void g2_encode_affine(const Fp2& x, const Fp2& y, bool infinity, uint8_t* rec) {
    if (infinity) {
        for (int i = 0; i < 129; i++) rec[i] = 0;
        return;
    }
    rec[0] = 4;
    fq2_encode(x, rec + 1);
    fq2_encode(y, rec + 65);
}

}  // namespace bn
