/*
 * bn_ref.c -- CPU restatement of the `bn` crate's BN254 path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the limb-level oracle and the timed CPU baseline ("port" of the reference
 * algorithm: same formulas, same Montgomery form, same operation order, including the
 * literal multiplications by -1 and by xi that the reference performs).  It is NOT the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Nothing under bn_b200/ links or calls it.
 *
 * Parity pin: tests/test_oracle.py checks it against the reference's known-answer tests
 * (tests/golden/, extracted from the reference's own test functions) and against the
 * independent big-integer oracle oracle/bn_oracle.py.
 *
 * Every function cites the reference file:line it restates (reference checkout paths).
 * Byte layouts == the crate's #[repr(C)] types: Montgomery form, canonical in [0,p),
 * four little-endian u64 limbs per field element.
 */
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fq_t;            /* src/arith.rs:9-11, src/fields/fp.rs:11-13 */
typedef struct { fq_t c0, c1; } fq2_t;             /* src/fields/fq2.rs:24-29 */
typedef struct { fq2_t c0, c1, c2; } fq6_t;        /* src/fields/fq6.rs:42-48 */
typedef struct { fq6_t c0, c1; } fq12_t;           /* src/fields/fq12.rs:26-31 */
typedef struct { fq_t x, y, z; } g1_t;             /* src/groups/mod.rs:36-41 */
typedef struct { fq2_t x, y, z; } g2_t;
typedef struct { fq2_t ell_0, ell_vw, ell_vv; } ell_t; /* src/groups/mod.rs:472-477 */

#include "bn_ref_consts.h"

/* ------------------------------------------------------------------ U256 (src/arith.rs) */
static int u256_cmp(const uint64_t *a, const uint64_t *b) { /* :161-174 */
    for (int i = 3; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
static int u256_is_zero(const uint64_t *a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static void add_nocarry(uint64_t *a, const uint64_t *b) { /* :408-416 */
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; a[i] = (uint64_t)c; c >>= 64; }
}
static void sub_noborrow(uint64_t *a, const uint64_t *b) { /* :419-439 */
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        a[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
    }
}
static void u256_add(uint64_t *a, const uint64_t *b, const uint64_t *m) { /* :238-244 */
    add_nocarry(a, b);
    if (u256_cmp(a, m) >= 0) sub_noborrow(a, m);
}
static void u256_sub(uint64_t *a, const uint64_t *b, const uint64_t *m) { /* :247-253 */
    if (u256_cmp(a, b) < 0) add_nocarry(a, m);
    sub_noborrow(a, b);
}
static void u256_neg(uint64_t *a, const uint64_t *m) { /* :266-273 */
    if (!u256_is_zero(a)) { uint64_t t[4]; memcpy(t, m, 32); sub_noborrow(t, a); memcpy(a, t, 32); }
}
/* acc[0..n) += b[0..4) * c, carry rippling through the remaining limbs (:441-478) */
static void mac_digit(uint64_t *acc, int n, const uint64_t *b, uint64_t c) {
    if (c == 0) return;
    uint64_t carry = 0;
    for (int i = 0; i < n; i++) {
        if (i < 4) {
            u128 t = (u128)b[i] * c + acc[i] + carry;
            acc[i] = (uint64_t)t; carry = (uint64_t)(t >> 64);
        } else if (carry) {
            u128 t = (u128)acc[i] + carry;
            acc[i] = (uint64_t)t; carry = (uint64_t)(t >> 64);
        } else break;
    }
}
/* Montgomery multiply, HAC 14.32 in separated-operand form (:481-503) + final correction (:257-263) */
static void u256_mul(uint64_t *a, const uint64_t *b, const uint64_t *m, uint64_t inv) {
    uint64_t res[8] = {0};
    for (int i = 0; i < 4; i++) mac_digit(res + i, 8 - i, b, a[i]);
    for (int i = 0; i < 4; i++) { uint64_t k = inv * res[i]; mac_digit(res + i, 8 - i, m, k); }
    memcpy(a, res + 4, 32);
    if (u256_cmp(a, m) >= 0) sub_noborrow(a, m);
}
static void div2(uint64_t *a) { /* :361-372 */
    a[0] = (a[0] >> 1) | (a[1] << 63); a[1] = (a[1] >> 1) | (a[2] << 63);
    a[2] = (a[2] >> 1) | (a[3] << 63); a[3] >>= 1;
}
/* binary extended Euclid, Guajardo et al. Alg. 16 (:281-327) */
static void u256_invert(uint64_t *a, const uint64_t *m) {
    static const uint64_t ONE[4] = {1, 0, 0, 0};
    uint64_t u[4], v[4], b[4] = {1, 0, 0, 0}, c[4] = {0, 0, 0, 0};
    memcpy(u, a, 32); memcpy(v, m, 32);
    while (u256_cmp(u, ONE) != 0 && u256_cmp(v, ONE) != 0) {
        while (!(u[0] & 1)) { div2(u); if (b[0] & 1) add_nocarry(b, m); div2(b); }
        while (!(v[0] & 1)) { div2(v); if (c[0] & 1) add_nocarry(c, m); div2(c); }
        if (u256_cmp(u, v) >= 0) { sub_noborrow(u, v); u256_sub(b, c, m); }
        else { sub_noborrow(v, u); u256_sub(c, b, m); }
    }
    memcpy(a, u256_cmp(u, ONE) == 0 ? b : c, 32);
}
static int u256_bit(const uint64_t *a, int n) { return (int)((a[n >> 6] >> (n & 63)) & 1); } /* :225-235 */

/* ------------------------------------------------------------------ Fq / Fr (src/fields/fp.rs) */
static const fq_t FQ_ZERO_C = {{0, 0, 0, 0}};
static fq_t fq_one(void) { fq_t r; memcpy(r.l, FQ_R1, 32); return r; }                 /* :90-92 */
static int fq_is_zero(fq_t a) { return u256_is_zero(a.l); }
static int fq_eq(fq_t a, fq_t b) { return memcmp(a.l, b.l, 32) == 0; }
static fq_t fq_add(fq_t a, fq_t b) { u256_add(a.l, b.l, FQ_MODULUS); return a; }        /* :115-124 */
static fq_t fq_sub(fq_t a, fq_t b) { u256_sub(a.l, b.l, FQ_MODULUS); return a; }        /* :126-135 */
static fq_t fq_mul(fq_t a, fq_t b) { u256_mul(a.l, b.l, FQ_MODULUS, FQ_INV); return a; } /* :137-146 */
static fq_t fq_neg(fq_t a) { u256_neg(a.l, FQ_MODULUS); return a; }                     /* :148-157 */
static fq_t fq_sqr(fq_t a) { return fq_mul(a, a); }                                     /* fields/mod.rs:31-33 */
static fq_t fq_inv(fq_t a) { /* :103-112 ; caller guarantees a != 0 */
    u256_invert(a.l, FQ_MODULUS); u256_mul(a.l, FQ_R3, FQ_MODULUS, FQ_INV); return a;
}
/* Fr: only what the path needs (scalar de-Montgomerization :15-22, mul/add/inverse for the vectors) */
static void fr_to_u256(const uint64_t *fr, uint64_t *out) {
    static const uint64_t ONE[4] = {1, 0, 0, 0};
    memcpy(out, fr, 32); u256_mul(out, ONE, FR_MODULUS, FR_INV);
}

/* ------------------------------------------------------------------ Fq2 (src/fields/fq2.rs) */
static fq2_t fq2_zero(void) { fq2_t r; r.c0 = FQ_ZERO_C; r.c1 = FQ_ZERO_C; return r; }
static fq2_t fq2_one(void) { fq2_t r; r.c0 = fq_one(); r.c1 = FQ_ZERO_C; return r; }
static int fq2_is_zero(fq2_t a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
static int fq2_eq(fq2_t a, fq2_t b) { return fq_eq(a.c0, b.c0) && fq_eq(a.c1, b.c1); }
static fq2_t fq2_add(fq2_t a, fq2_t b) { fq2_t r = {fq_add(a.c0, b.c0), fq_add(a.c1, b.c1)}; return r; }
static fq2_t fq2_sub(fq2_t a, fq2_t b) { fq2_t r = {fq_sub(a.c0, b.c0), fq_sub(a.c1, b.c1)}; return r; }
static fq2_t fq2_neg(fq2_t a) { fq2_t r = {fq_neg(a.c0), fq_neg(a.c1)}; return r; }
static fq2_t fq2_mul(fq2_t a, fq2_t b) { /* :139-155 */
    fq_t aa = fq_mul(a.c0, b.c0), bb = fq_mul(a.c1, b.c1);
    fq2_t r;
    r.c0 = fq_add(fq_mul(bb, FQ_NON_RESIDUE), aa);
    r.c1 = fq_sub(fq_sub(fq_mul(fq_add(a.c0, a.c1), fq_add(b.c0, b.c1)), aa), bb);
    return r;
}
static fq2_t fq2_sqr(fq2_t a) { /* :112-123 */
    fq_t ab = fq_mul(a.c0, a.c1);
    fq2_t r;
    r.c0 = fq_sub(fq_sub(fq_mul(fq_add(fq_mul(a.c1, FQ_NON_RESIDUE), a.c0), fq_add(a.c0, a.c1)), ab),
                  fq_mul(ab, FQ_NON_RESIDUE));
    r.c1 = fq_add(ab, ab);
    return r;
}
static fq2_t fq2_scale(fq2_t a, fq_t by) { fq2_t r = {fq_mul(a.c0, by), fq_mul(a.c1, by)}; return r; } /* :63-68 */
static fq2_t fq2_mul_by_nonresidue(fq2_t a) { return fq2_mul(a, FQ2_NONRESIDUE); }                     /* :70-72 */
static fq2_t fq2_frobenius_map(fq2_t a, int power) { /* :74-83 */
    if (power % 2 == 0) return a;
    fq2_t r = {a.c0, fq_mul(a.c1, FQ_NON_RESIDUE)};
    return r;
}
static fq2_t fq2_inv(fq2_t a) { /* :125-136 */
    fq_t t = fq_inv(fq_sub(fq_sqr(a.c0), fq_mul(fq_sqr(a.c1), FQ_NON_RESIDUE)));
    fq2_t r = {fq_mul(a.c0, t), fq_neg(fq_mul(a.c1, t))};
    return r;
}

/* ------------------------------------------------------------------ Fq6 (src/fields/fq6.rs) */
static fq6_t fq6_zero(void) { fq6_t r = {fq2_zero(), fq2_zero(), fq2_zero()}; return r; }
static fq6_t fq6_one(void) { fq6_t r = {fq2_one(), fq2_zero(), fq2_zero()}; return r; }
static fq6_t fq6_add(fq6_t a, fq6_t b) { fq6_t r = {fq2_add(a.c0, b.c0), fq2_add(a.c1, b.c1), fq2_add(a.c2, b.c2)}; return r; }
static fq6_t fq6_sub(fq6_t a, fq6_t b) { fq6_t r = {fq2_sub(a.c0, b.c0), fq2_sub(a.c1, b.c1), fq2_sub(a.c2, b.c2)}; return r; }
static fq6_t fq6_neg(fq6_t a) { fq6_t r = {fq2_neg(a.c0), fq2_neg(a.c1), fq2_neg(a.c2)}; return r; }
static fq6_t fq6_mul_by_nonresidue(fq6_t a) { fq6_t r = {fq2_mul_by_nonresidue(a.c2), a.c0, a.c1}; return r; } /* :59-65 */
static fq6_t fq6_scale(fq6_t a, fq2_t by) { fq6_t r = {fq2_mul(a.c0, by), fq2_mul(a.c1, by), fq2_mul(a.c2, by)}; return r; } /* :67-73 */
static fq6_t fq6_mul(fq6_t a, fq6_t b) { /* :144-158 */
    fq2_t a_a = fq2_mul(a.c0, b.c0), b_b = fq2_mul(a.c1, b.c1), c_c = fq2_mul(a.c2, b.c2);
    fq6_t r;
    r.c0 = fq2_add(fq2_mul_by_nonresidue(fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c1, a.c2), fq2_add(b.c1, b.c2)), b_b), c_c)), a_a);
    r.c1 = fq2_add(fq2_sub(fq2_sub(fq2_mul(fq2_add(a.c0, a.c1), fq2_add(b.c0, b.c1)), a_a), b_b), fq2_mul_by_nonresidue(c_c));
    r.c2 = fq2_sub(fq2_add(fq2_sub(fq2_mul(fq2_add(a.c0, a.c2), fq2_add(b.c0, b.c2)), a_a), b_b), c_c);
    return r;
}
static fq6_t fq6_sqr(fq6_t a) { /* :113-127 */
    fq2_t s0 = fq2_sqr(a.c0), ab = fq2_mul(a.c0, a.c1), s1 = fq2_add(ab, ab);
    fq2_t s2 = fq2_sqr(fq2_add(fq2_sub(a.c0, a.c1), a.c2));
    fq2_t bc = fq2_mul(a.c1, a.c2), s3 = fq2_add(bc, bc), s4 = fq2_sqr(a.c2);
    fq6_t r;
    r.c0 = fq2_add(s0, fq2_mul_by_nonresidue(s3));
    r.c1 = fq2_add(s1, fq2_mul_by_nonresidue(s4));
    r.c2 = fq2_sub(fq2_sub(fq2_add(fq2_add(s1, s2), s3), s0), s4);
    return r;
}
static fq6_t fq6_inv(fq6_t a) { /* :129-141 */
    fq2_t c0 = fq2_sub(fq2_sqr(a.c0), fq2_mul(a.c1, fq2_mul_by_nonresidue(a.c2)));
    fq2_t c1 = fq2_sub(fq2_mul_by_nonresidue(fq2_sqr(a.c2)), fq2_mul(a.c0, a.c1));
    fq2_t c2 = fq2_sub(fq2_sqr(a.c1), fq2_mul(a.c0, a.c2));
    fq2_t t = fq2_inv(fq2_add(fq2_mul_by_nonresidue(fq2_add(fq2_mul(a.c2, c1), fq2_mul(a.c1, c2))), fq2_mul(a.c0, c0)));
    fq6_t r = {fq2_mul(t, c0), fq2_mul(t, c1), fq2_mul(t, c2)};
    return r;
}
static fq6_t fq6_frobenius_map(fq6_t a, int power) { /* :75-81 */
    fq6_t r;
    r.c0 = fq2_frobenius_map(a.c0, power);
    r.c1 = fq2_mul(fq2_frobenius_map(a.c1, power), FQ6_FROB_C1[power]);
    r.c2 = fq2_mul(fq2_frobenius_map(a.c2, power), FQ6_FROB_C2[power]);
    return r;
}

/* ------------------------------------------------------------------ Fq12 (src/fields/fq12.rs) */
static fq12_t fq12_one(void) { fq12_t r = {fq6_one(), fq6_zero()}; return r; }
static fq12_t fq12_add(fq12_t a, fq12_t b) { fq12_t r = {fq6_add(a.c0, b.c0), fq6_add(a.c1, b.c1)}; return r; }
static fq12_t fq12_sub(fq12_t a, fq12_t b) { fq12_t r = {fq6_sub(a.c0, b.c0), fq6_sub(a.c1, b.c1)}; return r; }
static fq12_t fq12_neg(fq12_t a) { fq12_t r = {fq6_neg(a.c0), fq6_neg(a.c1)}; return r; }
static fq12_t fq12_mul(fq12_t a, fq12_t b) { /* :295-307 */
    fq6_t aa = fq6_mul(a.c0, b.c0), bb = fq6_mul(a.c1, b.c1);
    fq12_t r;
    r.c0 = fq6_add(fq6_mul_by_nonresidue(bb), aa);
    r.c1 = fq6_sub(fq6_sub(fq6_mul(fq6_add(a.c0, a.c1), fq6_add(b.c0, b.c1)), aa), bb);
    return r;
}
static fq12_t fq12_sqr(fq12_t a) { /* :275-282 */
    fq6_t ab = fq6_mul(a.c0, a.c1);
    fq12_t r;
    r.c0 = fq6_sub(fq6_sub(fq6_mul(fq6_add(fq6_mul_by_nonresidue(a.c1), a.c0), fq6_add(a.c0, a.c1)), ab), fq6_mul_by_nonresidue(ab));
    r.c1 = fq6_add(ab, ab);
    return r;
}
static fq12_t fq12_inv(fq12_t a) { /* :284-292 */
    fq6_t t = fq6_inv(fq6_sub(fq6_sqr(a.c0), fq6_mul_by_nonresidue(fq6_sqr(a.c1))));
    fq12_t r = {fq6_mul(a.c0, t), fq6_neg(fq6_mul(a.c1, t))};
    return r;
}
static fq12_t fq12_unitary_inverse(fq12_t a) { fq12_t r = {a.c0, fq6_neg(a.c1)}; return r; } /* :103-105 */
static fq12_t fq12_frobenius_map(fq12_t a, int power) { /* :90-95 */
    fq12_t r = {fq6_frobenius_map(a.c0, power), fq6_scale(fq6_frobenius_map(a.c1, power), FQ12_FROB_C1[power])};
    return r;
}
static fq12_t fq12_mul_by_024(fq12_t a, fq2_t ell_0, fq2_t ell_vw, fq2_t ell_vv) { /* :107-176 */
    fq2_t z0 = a.c0.c0, z1 = a.c0.c1, z2 = a.c0.c2, z3 = a.c1.c0, z4 = a.c1.c1, z5 = a.c1.c2;
    fq2_t x0 = ell_0, x2 = ell_vv, x4 = ell_vw;
    fq2_t d0 = fq2_mul(z0, x0), d2 = fq2_mul(z2, x2), d4 = fq2_mul(z4, x4);
    fq2_t t2 = fq2_add(z0, z4), t1 = fq2_add(z0, z2), s0 = fq2_add(fq2_add(z1, z3), z5);
    fq2_t s1 = fq2_mul(z1, x2), t3 = fq2_add(s1, d4), t4 = fq2_add(fq2_mul_by_nonresidue(t3), d0), t0;
    z0 = t4;
    t3 = fq2_mul(z5, x4); s1 = fq2_add(s1, t3); t3 = fq2_add(t3, d2); t4 = fq2_mul_by_nonresidue(t3);
    t3 = fq2_mul(z1, x0); s1 = fq2_add(s1, t3); t4 = fq2_add(t4, t3);
    z1 = t4;
    t0 = fq2_add(x0, x2); t3 = fq2_sub(fq2_sub(fq2_mul(t1, t0), d0), d2);
    t4 = fq2_mul(z3, x4); s1 = fq2_add(s1, t4); t3 = fq2_add(t3, t4);
    t0 = fq2_add(z2, z4);
    z2 = t3;
    t1 = fq2_add(x2, x4); t3 = fq2_sub(fq2_sub(fq2_mul(t0, t1), d2), d4); t4 = fq2_mul_by_nonresidue(t3);
    t3 = fq2_mul(z3, x0); s1 = fq2_add(s1, t3); t4 = fq2_add(t4, t3);
    z3 = t4;
    t3 = fq2_mul(z5, x2); s1 = fq2_add(s1, t3); t4 = fq2_mul_by_nonresidue(t3);
    t0 = fq2_add(x0, x4); t3 = fq2_sub(fq2_sub(fq2_mul(t2, t0), d0), d4); t4 = fq2_add(t4, t3);
    z4 = t4;
    t0 = fq2_add(fq2_add(x0, x2), x4); t3 = fq2_sub(fq2_mul(s0, t0), s1);
    z5 = t3;
    fq12_t r = {{z0, z1, z2}, {z3, z4, z5}};
    return r;
}
static void fp4_sqr_gs(fq2_t x, fq2_t y, fq2_t *re, fq2_t *im) { /* the repeated block at :186-196 */
    fq2_t tmp = fq2_mul(x, y);
    *re = fq2_sub(fq2_sub(fq2_mul(fq2_add(x, y), fq2_add(fq2_mul_by_nonresidue(y), x)), tmp), fq2_mul_by_nonresidue(tmp));
    *im = fq2_add(tmp, tmp);
}
static fq12_t fq12_cyclotomic_squared(fq12_t a) { /* :178-227 */
    fq2_t z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2;
    fq2_t t0, t1, t2, t3, t4, t5, tmp;
    fp4_sqr_gs(z0, z1, &t0, &t1);
    fp4_sqr_gs(z2, z3, &t2, &t3);
    fp4_sqr_gs(z4, z5, &t4, &t5);
    z0 = fq2_sub(t0, z0); z0 = fq2_add(z0, z0); z0 = fq2_add(z0, t0);
    z1 = fq2_add(t1, z1); z1 = fq2_add(z1, z1); z1 = fq2_add(z1, t1);
    tmp = fq2_mul_by_nonresidue(t5);
    z2 = fq2_add(tmp, z2); z2 = fq2_add(z2, z2); z2 = fq2_add(z2, tmp);
    z3 = fq2_sub(t4, z3); z3 = fq2_add(z3, z3); z3 = fq2_add(z3, t4);
    z4 = fq2_sub(t2, z4); z4 = fq2_add(z4, z4); z4 = fq2_add(z4, t2);
    z5 = fq2_add(t3, z5); z5 = fq2_add(z5, z5); z5 = fq2_add(z5, t3);
    fq12_t r = {{z0, z4, z3}, {z2, z1, z5}};
    return r;
}
static fq12_t fq12_cyclotomic_pow(fq12_t a, const uint64_t *by) { /* :229-246 */
    fq12_t res = fq12_one();
    int found_one = 0;
    for (int i = 255; i >= 0; i--) {
        if (found_one) res = fq12_cyclotomic_squared(res);
        if (u256_bit(by, i)) { found_one = 1; res = fq12_mul(a, res); }
    }
    return res;
}
static fq12_t fq12_exp_by_neg_z(fq12_t a) { return fq12_unitary_inverse(fq12_cyclotomic_pow(a, BN_U)); } /* :97-101 */
static fq12_t fq12_pow(fq12_t a, const uint64_t *by) { /* fields/mod.rs:35-46 */
    fq12_t res = fq12_one();
    for (int i = 255; i >= 0; i--) {
        res = fq12_sqr(res);
        if (u256_bit(by, i)) res = fq12_mul(a, res);
    }
    return res;
}
static fq12_t final_exponentiation_first_chunk(fq12_t s) { /* :41-52 */
    fq12_t b = fq12_inv(s), a = fq12_unitary_inverse(s), c = fq12_mul(a, b), d = fq12_frobenius_map(c, 2);
    return fq12_mul(d, c);
}
static fq12_t final_exponentiation_last_chunk(fq12_t s) { /* :54-84 */
    fq12_t a = fq12_exp_by_neg_z(s), b = fq12_cyclotomic_squared(a), c = fq12_cyclotomic_squared(b), d = fq12_mul(c, b);
    fq12_t e = fq12_exp_by_neg_z(d), f = fq12_cyclotomic_squared(e), g = fq12_exp_by_neg_z(f);
    fq12_t h = fq12_unitary_inverse(d), i = fq12_unitary_inverse(g);
    fq12_t j = fq12_mul(i, e), k = fq12_mul(j, h), l = fq12_mul(k, b), m = fq12_mul(k, e), n = fq12_mul(s, m);
    fq12_t o = fq12_frobenius_map(l, 1), p = fq12_mul(o, n), q = fq12_frobenius_map(k, 2), r = fq12_mul(q, p);
    fq12_t ss = fq12_unitary_inverse(s), t = fq12_mul(ss, l), u = fq12_frobenius_map(t, 3);
    return fq12_mul(u, r);
}
static fq12_t final_exponentiation(fq12_t f) { return final_exponentiation_last_chunk(final_exponentiation_first_chunk(f)); } /* :86-88 */

/* ------------------------------------------------------------------ groups (src/groups/mod.rs) */
/* The reference is generic over P::Base; C has no generics, so the law is stated twice via a macro. */
#define DEFINE_GROUP(G, F, PFX)                                                                        \
    static G PFX##_zero(void) { G r = {F##_zero_v(), F##_one_v(), F##_zero_v()}; return r; } /* :208-214 */ \
    static int PFX##_is_zero(G p) { return F##_is_zero(p.z); }                              /* :224-226 */ \
    static G PFX##_double(G p) { /* :228-247 */                                                        \
        F##_t a = F##_sqr(p.x), b = F##_sqr(p.y), c = F##_sqr(b);                                      \
        F##_t d = F##_sub(F##_sub(F##_sqr(F##_add(p.x, b)), a), c);                                    \
        d = F##_add(d, d);                                                                             \
        F##_t e = F##_add(F##_add(a, a), a), f = F##_sqr(e);                                           \
        F##_t x3 = F##_sub(f, F##_add(d, d));                                                          \
        F##_t eight_c = F##_add(c, c);                                                                 \
        eight_c = F##_add(eight_c, eight_c); eight_c = F##_add(eight_c, eight_c);                      \
        F##_t y1z1 = F##_mul(p.y, p.z);                                                                \
        G r = {x3, F##_sub(F##_mul(e, F##_sub(d, x3)), eight_c), F##_add(y1z1, y1z1)};                 \
        return r;                                                                                      \
    }                                                                                                  \
    static G PFX##_add(G p, G o) { /* :272-312 */                                                      \
        if (PFX##_is_zero(p)) return o;                                                                \
        if (PFX##_is_zero(o)) return p;                                                                \
        F##_t z1s = F##_sqr(p.z), z2s = F##_sqr(o.z);                                                  \
        F##_t u1 = F##_mul(p.x, z2s), u2 = F##_mul(o.x, z1s);                                          \
        F##_t z1c = F##_mul(p.z, z1s), z2c = F##_mul(o.z, z2s);                                        \
        F##_t s1 = F##_mul(p.y, z2c), s2 = F##_mul(o.y, z1c);                                          \
        if (F##_eq(u1, u2) && F##_eq(s1, s2)) return PFX##_double(p);                                  \
        F##_t h = F##_sub(u2, u1), s2ms1 = F##_sub(s2, s1);                                            \
        F##_t i = F##_sqr(F##_add(h, h)), j = F##_mul(h, i), r_ = F##_add(s2ms1, s2ms1);               \
        F##_t v = F##_mul(u1, i), s1j = F##_mul(s1, j);                                                \
        F##_t x3 = F##_sub(F##_sub(F##_sqr(r_), j), F##_add(v, v));                                    \
        G r = {x3, F##_sub(F##_mul(r_, F##_sub(v, x3)), F##_add(s1j, s1j)),                            \
               F##_mul(F##_sub(F##_sub(F##_sqr(F##_add(p.z, o.z)), z1s), z2s), h)};                    \
        return r;                                                                                      \
    }                                                                                                  \
    static G PFX##_neg(G p) { /* :314-328 */                                                           \
        if (PFX##_is_zero(p)) return p;                                                                \
        p.y = F##_neg(p.y); return p;                                                                  \
    }                                                                                                  \
    static G PFX##_mul(G p, const uint64_t *fr_mont) { /* :250-270 */                                  \
        uint64_t k[4]; fr_to_u256(fr_mont, k);                                                         \
        G res = PFX##_zero(); int found_one = 0;                                                       \
        for (int i = 255; i >= 0; i--) {                                                               \
            if (found_one) res = PFX##_double(res);                                                    \
            if (u256_bit(k, i)) { found_one = 1; res = PFX##_add(res, p); }                            \
        }                                                                                              \
        return res;                                                                                    \
    }                                                                                                  \
    /* returns 0 for infinity (None), else 1 (:113-130) */                                             \
    static int PFX##_to_affine(G p, F##_t *x, F##_t *y) {                                              \
        if (F##_is_zero(p.z)) return 0;                                                                \
        if (F##_eq(p.z, F##_one_v())) { *x = p.x; *y = p.y; return 1; }                                \
        F##_t zinv = F##_inv(p.z), zinv2 = F##_sqr(zinv);                                              \
        *x = F##_mul(p.x, zinv2); *y = F##_mul(p.y, F##_mul(zinv2, zinv));                             \
        return 1;                                                                                      \
    }

#define fq_zero_v() (FQ_ZERO_C)
#define fq_one_v() (fq_one())
#define fq2_zero_v() (fq2_zero())
#define fq2_one_v() (fq2_one())
DEFINE_GROUP(g1_t, fq, g1)
DEFINE_GROUP(g2_t, fq2, g2)

/* ------------------------------------------------------------------ pairing (src/groups/mod.rs:441-634, 764-771) */
static ell_t mixed_addition_step(g2_t *r, fq2_t bx, fq2_t by) { /* :592-610 */
    fq2_t d = fq2_sub(r->x, fq2_mul(r->z, bx)), e = fq2_sub(r->y, fq2_mul(r->z, by));
    fq2_t f = fq2_sqr(d), g = fq2_sqr(e), h = fq2_mul(d, f), i = fq2_mul(r->x, f);
    fq2_t j = fq2_sub(fq2_add(fq2_mul(r->z, g), h), fq2_add(i, i));
    r->x = fq2_mul(d, j);
    r->y = fq2_sub(fq2_mul(e, fq2_sub(i, j)), fq2_mul(h, r->y));
    r->z = fq2_mul(r->z, h);
    ell_t c;
    c.ell_0 = fq2_mul(FQ2_NONRESIDUE, fq2_sub(fq2_mul(e, bx), fq2_mul(d, by)));
    c.ell_vv = fq2_neg(e);
    c.ell_vw = d;
    return c;
}
static ell_t doubling_step(g2_t *r) { /* :612-634 */
    fq2_t a = fq2_scale(fq2_mul(r->x, r->y), TWO_INV), b = fq2_sqr(r->y), c = fq2_sqr(r->z);
    fq2_t d = fq2_add(fq2_add(c, c), c), e = fq2_mul(G2_B, d), f = fq2_add(fq2_add(e, e), e);
    fq2_t g = fq2_scale(fq2_add(b, f), TWO_INV);
    fq2_t h = fq2_sub(fq2_sqr(fq2_add(r->y, r->z)), fq2_add(b, c));
    fq2_t i = fq2_sub(e, b), j = fq2_sqr(r->x), e_sq = fq2_sqr(e);
    r->x = fq2_mul(a, fq2_sub(b, f));
    r->y = fq2_sub(fq2_sqr(g), fq2_add(fq2_add(e_sq, e_sq), e_sq));
    r->z = fq2_mul(b, h);
    ell_t out;
    out.ell_0 = fq2_mul(FQ2_NONRESIDUE, i);
    out.ell_vw = fq2_neg(h);
    out.ell_vv = fq2_add(fq2_add(j, j), j);
    return out;
}
static void mul_by_q(fq2_t x, fq2_t y, fq2_t *ox, fq2_t *oy) { /* :550-555 */
    *ox = fq2_mul(TWIST_MUL_BY_Q_X, fq2_frobenius_map(x, 1));
    *oy = fq2_mul(TWIST_MUL_BY_Q_Y, fq2_frobenius_map(y, 1));
}
static int g2_precompute(fq2_t qx, fq2_t qy, ell_t *coeffs) { /* :557-588 ; returns count (102) */
    g2_t r = {qx, qy, fq2_one()};
    int n = 0, found_one = 0;
    for (int b = 255; b >= 0; b--) {
        int i = u256_bit(ATE_LOOP_COUNT, b);
        if (!found_one) { found_one = i; continue; }
        coeffs[n++] = doubling_step(&r);
        if (i) coeffs[n++] = mixed_addition_step(&r, qx, qy);
    }
    fq2_t q1x, q1y, q2x, q2y;
    mul_by_q(qx, qy, &q1x, &q1y);
    mul_by_q(q1x, q1y, &q2x, &q2y);
    q2y = fq2_neg(q2y);
    coeffs[n++] = mixed_addition_step(&r, q1x, q1y);
    coeffs[n++] = mixed_addition_step(&r, q2x, q2y);
    return n;
}
static fq12_t miller_loop(const ell_t *coeffs, fq_t px, fq_t py) { /* :485-520 */
    fq12_t f = fq12_one();
    int idx = 0, found_one = 0;
    for (int b = 255; b >= 0; b--) {
        int i = u256_bit(ATE_LOOP_COUNT, b);
        if (!found_one) { found_one = i; continue; }
        const ell_t *c = &coeffs[idx++];
        f = fq12_mul_by_024(fq12_sqr(f), c->ell_0, fq2_scale(c->ell_vw, py), fq2_scale(c->ell_vv, px));
        if (i) {
            c = &coeffs[idx++];
            f = fq12_mul_by_024(f, c->ell_0, fq2_scale(c->ell_vw, py), fq2_scale(c->ell_vv, px));
        }
    }
    for (int k = 0; k < 2; k++) {
        const ell_t *c = &coeffs[idx++];
        f = fq12_mul_by_024(f, c->ell_0, fq2_scale(c->ell_vw, py), fq2_scale(c->ell_vv, px));
    }
    return f;
}
static fq12_t pairing(g1_t p, g2_t q) { /* :764-771 */
    fq_t px, py; fq2_t qx, qy;
    if (!g1_to_affine(p, &px, &py) || !g2_to_affine(q, &qx, &qy)) return fq12_one();
    ell_t coeffs[102];
    g2_precompute(qx, qy, coeffs);
    return final_exponentiation(miller_loop(coeffs, px, py));
}

/* ------------------------------------------------------------------ exported batch API */
typedef void (*item_fn)(size_t i, void *ctx);
typedef struct { item_fn fn; void *ctx; size_t lo, hi; } job_t;
static void *job_main(void *p) { job_t *j = (job_t *)p; for (size_t i = j->lo; i < j->hi; i++) j->fn(i, j->ctx); return NULL; }
static void par_for(size_t n, int threads, item_fn fn, void *ctx) {
    if (threads <= 1 || n < 2) { for (size_t i = 0; i < n; i++) fn(i, ctx); return; }
    if ((size_t)threads > n) threads = (int)n;
    pthread_t *tid = malloc(sizeof(pthread_t) * threads);
    job_t *jobs = malloc(sizeof(job_t) * threads);
    for (int t = 0; t < threads; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
        pthread_create(&tid[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
}

typedef struct { const void *a, *b; void *out; uint32_t iters; } bctx_t;
#define EXPORT __attribute__((visibility("default")))

static void pairing_item(size_t i, void *c) { bctx_t *x = c; ((fq12_t *)x->out)[i] = pairing(((const g1_t *)x->a)[i], ((const g2_t *)x->b)[i]); }
EXPORT void bn_ref_pairing_batch(const g1_t *p, const g2_t *q, fq12_t *out, size_t n, int threads) {
    bctx_t c = {p, q, out, 0}; par_for(n, threads, pairing_item, &c);
}
static void g1mul_item(size_t i, void *c) { bctx_t *x = c; ((g1_t *)x->out)[i] = g1_mul(((const g1_t *)x->a)[i], ((const fq_t *)x->b)[i].l); }
EXPORT void bn_ref_g1_mul_batch(const g1_t *p, const fq_t *k, g1_t *out, size_t n, int threads) {
    bctx_t c = {p, k, out, 0}; par_for(n, threads, g1mul_item, &c);
}
static void g2mul_item(size_t i, void *c) { bctx_t *x = c; ((g2_t *)x->out)[i] = g2_mul(((const g2_t *)x->a)[i], ((const fq_t *)x->b)[i].l); }
EXPORT void bn_ref_g2_mul_batch(const g2_t *p, const fq_t *k, g2_t *out, size_t n, int threads) {
    bctx_t c = {p, k, out, 0}; par_for(n, threads, g2mul_item, &c);
}
static void gtpow_item(size_t i, void *c) {
    bctx_t *x = c; uint64_t k[4]; fr_to_u256(((const fq_t *)x->b)[i].l, k);          /* lib.rs:171 */
    ((fq12_t *)x->out)[i] = fq12_pow(((const fq12_t *)x->a)[i], k);
}
EXPORT void bn_ref_gt_pow_batch(const fq12_t *a, const fq_t *k, fq12_t *out, size_t n, int threads) {
    bctx_t c = {a, k, out, 0}; par_for(n, threads, gtpow_item, &c);
}
static void gtmul_item(size_t i, void *c) { bctx_t *x = c; ((fq12_t *)x->out)[i] = fq12_mul(((const fq12_t *)x->a)[i], ((const fq12_t *)x->b)[i]); }
EXPORT void bn_ref_gt_mul_batch(const fq12_t *a, const fq12_t *b, fq12_t *out, size_t n, int threads) {
    bctx_t c = {a, b, out, 0}; par_for(n, threads, gtmul_item, &c);
}
/* x <- x*b repeated `iters` times per lane: CPU counterpart of BASELINE config 2 */
static void fqchain_item(size_t i, void *c) {
    bctx_t *x = c; fq_t v = ((const fq_t *)x->a)[i], b = ((const fq_t *)x->b)[i];
    for (uint32_t k = 0; k < x->iters; k++) v = fq_mul(v, b);
    ((fq_t *)x->out)[i] = v;
}
EXPORT void bn_ref_fq_mul_chain(const fq_t *a, const fq_t *b, fq_t *out, size_t n, uint32_t iters, int threads) {
    bctx_t c = {a, b, out, iters}; par_for(n, threads, fqchain_item, &c);
}

/* ---- single-shot entry points used only by the oracle's own parity tests ---- */
EXPORT void bn_ref_g1_add(const g1_t *a, const g1_t *b, g1_t *out) { *out = g1_add(*a, *b); }
EXPORT void bn_ref_g2_add(const g2_t *a, const g2_t *b, g2_t *out) { *out = g2_add(*a, *b); }
EXPORT void bn_ref_g1_double(const g1_t *a, g1_t *out) { *out = g1_double(*a); }
EXPORT void bn_ref_g2_double(const g2_t *a, g2_t *out) { *out = g2_double(*a); }
EXPORT void bn_ref_g1_neg(const g1_t *a, g1_t *out) { *out = g1_neg(*a); }
EXPORT void bn_ref_g2_neg(const g2_t *a, g2_t *out) { *out = g2_neg(*a); }
EXPORT void bn_ref_g1_generator(g1_t *out) { out->x = G1_GEN_X; out->y = G1_GEN_Y; out->z = fq_one(); }
EXPORT void bn_ref_g2_generator(g2_t *out) { out->x = G2_GEN_X; out->y = G2_GEN_Y; out->z = fq2_one(); }
/* normalize (lib.rs:88-95): returns 0 and leaves *out = *a for infinity */
EXPORT int bn_ref_g1_normalize(const g1_t *a, g1_t *out) {
    *out = *a; fq_t x, y; if (!g1_to_affine(*a, &x, &y)) return 0;
    out->x = x; out->y = y; out->z = fq_one(); return 1;
}
EXPORT int bn_ref_g2_normalize(const g2_t *a, g2_t *out) {
    *out = *a; fq2_t x, y; if (!g2_to_affine(*a, &x, &y)) return 0;
    out->x = x; out->y = y; out->z = fq2_one(); return 1;
}
EXPORT void bn_ref_fq12_mul(const fq12_t *a, const fq12_t *b, fq12_t *out) { *out = fq12_mul(*a, *b); }
EXPORT void bn_ref_fq12_sqr(const fq12_t *a, fq12_t *out) { *out = fq12_sqr(*a); }
EXPORT void bn_ref_fq12_add(const fq12_t *a, const fq12_t *b, fq12_t *out) { *out = fq12_add(*a, *b); }
EXPORT void bn_ref_fq12_sub(const fq12_t *a, const fq12_t *b, fq12_t *out) { *out = fq12_sub(*a, *b); }
EXPORT void bn_ref_fq12_neg(const fq12_t *a, fq12_t *out) { *out = fq12_neg(*a); }
EXPORT void bn_ref_fq12_inv(const fq12_t *a, fq12_t *out) { *out = fq12_inv(*a); }
EXPORT void bn_ref_fq12_exp_by_neg_z(const fq12_t *a, fq12_t *out) { *out = fq12_exp_by_neg_z(*a); }
EXPORT void bn_ref_fq12_frobenius(const fq12_t *a, int power, fq12_t *out) { *out = fq12_frobenius_map(*a, power); }
EXPORT void bn_ref_final_exponentiation(const fq12_t *a, fq12_t *out) { *out = final_exponentiation(*a); }
/* precompute + miller loop on AFFINE inputs (pins test_prepared_g2 / test_miller_loop) */
EXPORT int bn_ref_g2_precompute(const fq2_t *qx, const fq2_t *qy, ell_t *coeffs102) { return g2_precompute(*qx, *qy, coeffs102); }
EXPORT void bn_ref_miller_loop(const ell_t *coeffs102, const fq_t *px, const fq_t *py, fq12_t *out) { *out = miller_loop(coeffs102, *px, *py); }
/* generic modular helpers for Fr vectors: which = 0 -> Fq, 1 -> Fr */
EXPORT void bn_ref_fp_mul(int which, const fq_t *a, const fq_t *b, fq_t *out) {
    *out = *a; u256_mul(out->l, b->l, which ? FR_MODULUS : FQ_MODULUS, which ? FR_INV : FQ_INV);
}
EXPORT void bn_ref_fp_add(int which, const fq_t *a, const fq_t *b, fq_t *out) { *out = *a; u256_add(out->l, b->l, which ? FR_MODULUS : FQ_MODULUS); }
EXPORT void bn_ref_fp_sub(int which, const fq_t *a, const fq_t *b, fq_t *out) { *out = *a; u256_sub(out->l, b->l, which ? FR_MODULUS : FQ_MODULUS); }
EXPORT void bn_ref_fp_neg(int which, const fq_t *a, fq_t *out) { *out = *a; u256_neg(out->l, which ? FR_MODULUS : FQ_MODULUS); }
EXPORT int bn_ref_fp_inv(int which, const fq_t *a, fq_t *out) { /* fp.rs:103-112 */
    if (u256_is_zero(a->l)) return 0;
    *out = *a; u256_invert(out->l, which ? FR_MODULUS : FQ_MODULUS);
    u256_mul(out->l, which ? FR_R3 : FQ_R3, which ? FR_MODULUS : FQ_MODULUS, which ? FR_INV : FQ_INV);
    return 1;
}
EXPORT int bn_ref_abi_sizes(size_t *out6) {
    out6[0] = sizeof(fq_t); out6[1] = sizeof(fq2_t); out6[2] = sizeof(fq12_t);
    out6[3] = sizeof(g1_t); out6[4] = sizeof(g2_t); out6[5] = sizeof(ell_t);
    return 0;
}
