"""Big-integer CPU oracle for the BN254 path of the `bn` crate (TEST INFRASTRUCTURE ONLY).

This file is a plain-Python restatement of the reference algorithm.  It is the
checker, never the product: only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it.

Every value is a plain integer mod p ("plain domain", NOT Montgomery form);
`to_mont` / `from_mont` convert at the byte-image boundary.  Each function cites
the reference file:line it restates (paths relative to the reference checkout).

Parity pin: validated against the reference's own known-answer tests
(tests/golden/*.json, extracted by tests/golden/make_golden.py):
test_miller_loop, test_reduced_pairing, test_prepared_g2, fq12_test_vector,
test_cyclotomic_exp, test_str, g1/g2/fr serialization vectors.
"""
from __future__ import annotations

# --------------------------------------------------------------------------
# Curve constants (src/fields/fp.rs:161-177, src/fields/fq12.rs:99,
# src/groups/mod.rs:452-454)
# --------------------------------------------------------------------------
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R_ORDER = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
U = 4965661367192848881
ATE_LOOP_COUNT = 6 * U + 2
MONT_R = 1 << 256

assert Q == 36 * U**4 + 36 * U**3 + 24 * U**2 + 6 * U + 1
assert R_ORDER == 36 * U**4 + 36 * U**3 + 18 * U**2 + 6 * U + 1
assert ATE_LOOP_COUNT == 0x19D797039BE763BA8


def to_mont(x: int, p: int = Q) -> int:
    return (x * MONT_R) % p


def from_mont(x: int, p: int = Q) -> int:
    return (x * pow(MONT_R, -1, p)) % p


# --------------------------------------------------------------------------
# Fq (plain domain).  src/arith.rs:238-327, src/fields/fp.rs:103-157
# --------------------------------------------------------------------------
def fq_inv(a: int) -> int:
    assert a % Q != 0
    return pow(a, -1, Q)


# --------------------------------------------------------------------------
# Fq2 = Fq[i]/(i^2+1).  src/fields/fq2.rs
# --------------------------------------------------------------------------
XI = (9, 1)  # fq2_nonresidue(), src/fields/fq2.rs:17-22
FQ2_ZERO = (0, 0)
FQ2_ONE = (1, 0)


def fq2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def fq2_neg(a):
    return ((-a[0]) % Q, (-a[1]) % Q)


def fq2_mul(a, b):
    # src/fields/fq2.rs:139-155
    aa = a[0] * b[0]
    bb = a[1] * b[1]
    return ((aa - bb) % Q, ((a[0] + a[1]) * (b[0] + b[1]) - aa - bb) % Q)


def fq2_sqr(a):
    # src/fields/fq2.rs:112-123
    ab = a[0] * a[1]
    return (((a[0] - a[1]) * (a[0] + a[1])) % Q, (2 * ab) % Q)


def fq2_scale(a, k: int):
    # src/fields/fq2.rs:63-68
    return ((a[0] * k) % Q, (a[1] * k) % Q)


def fq2_mul_xi(a):
    # src/fields/fq2.rs:70-72
    return fq2_mul(a, XI)


def fq2_conj(a):
    return (a[0], (-a[1]) % Q)


def fq2_frob(a, power: int):
    # src/fields/fq2.rs:74-83
    return a if power % 2 == 0 else fq2_conj(a)


def fq2_inv(a):
    # src/fields/fq2.rs:125-136
    t = fq_inv((a[0] * a[0] + a[1] * a[1]) % Q)
    return ((a[0] * t) % Q, (-(a[1] * t)) % Q)


def fq2_pow(a, e: int):
    res = FQ2_ONE
    for bit in bin(e)[2:]:
        res = fq2_sqr(res)
        if bit == "1":
            res = fq2_mul(res, a)
    return res


# --------------------------------------------------------------------------
# Fq6 = Fq2[v]/(v^3 - xi).  src/fields/fq6.rs
# --------------------------------------------------------------------------
FQ6_ZERO = (FQ2_ZERO, FQ2_ZERO, FQ2_ZERO)
FQ6_ONE = (FQ2_ONE, FQ2_ZERO, FQ2_ZERO)


def _frob_gamma(power: int, k: int):
    """xi^(k*(q^power-1)/6): the Frobenius twist constants.
    k=1: src/fields/fq12.rs:7-24; k=2,4: src/fields/fq6.rs:5-40;
    (power=1,k=2),(power=1,k=3): src/groups/mod.rs:456-470."""
    return fq2_pow(XI, k * (Q**power - 1) // 6)


_GAMMA_CACHE: dict = {}


def frob_gamma(power: int, k: int):
    key = (power, k)
    if key not in _GAMMA_CACHE:
        _GAMMA_CACHE[key] = _frob_gamma(power, k)
    return _GAMMA_CACHE[key]


def fq6_add(a, b):
    return tuple(fq2_add(x, y) for x, y in zip(a, b))


def fq6_sub(a, b):
    return tuple(fq2_sub(x, y) for x, y in zip(a, b))


def fq6_neg(a):
    return tuple(fq2_neg(x) for x in a)


def fq6_mul_by_nonresidue(a):
    # src/fields/fq6.rs:59-65
    return (fq2_mul_xi(a[2]), a[0], a[1])


def fq6_mul(a, b):
    # src/fields/fq6.rs:144-158
    a_a = fq2_mul(a[0], b[0])
    b_b = fq2_mul(a[1], b[1])
    c_c = fq2_mul(a[2], b[2])
    c0 = fq2_add(
        fq2_mul_xi(fq2_sub(fq2_sub(fq2_mul(fq2_add(a[1], a[2]), fq2_add(b[1], b[2])), b_b), c_c)), a_a)
    c1 = fq2_add(
        fq2_sub(fq2_sub(fq2_mul(fq2_add(a[0], a[1]), fq2_add(b[0], b[1])), a_a), b_b), fq2_mul_xi(c_c))
    c2 = fq2_sub(
        fq2_add(fq2_sub(fq2_mul(fq2_add(a[0], a[2]), fq2_add(b[0], b[2])), a_a), b_b), c_c)
    return (c0, c1, c2)


def fq6_sqr(a):
    # src/fields/fq6.rs:113-127
    s0 = fq2_sqr(a[0])
    ab = fq2_mul(a[0], a[1])
    s1 = fq2_add(ab, ab)
    s2 = fq2_sqr(fq2_add(fq2_sub(a[0], a[1]), a[2]))
    bc = fq2_mul(a[1], a[2])
    s3 = fq2_add(bc, bc)
    s4 = fq2_sqr(a[2])
    return (
        fq2_add(s0, fq2_mul_xi(s3)),
        fq2_add(s1, fq2_mul_xi(s4)),
        fq2_sub(fq2_sub(fq2_add(fq2_add(s1, s2), s3), s0), s4),
    )


def fq6_inv(a):
    # src/fields/fq6.rs:129-141
    c0 = fq2_sub(fq2_sqr(a[0]), fq2_mul(a[1], fq2_mul_xi(a[2])))
    c1 = fq2_sub(fq2_mul_xi(fq2_sqr(a[2])), fq2_mul(a[0], a[1]))
    c2 = fq2_sub(fq2_sqr(a[1]), fq2_mul(a[0], a[2]))
    t = fq2_inv(fq2_add(fq2_mul_xi(fq2_add(fq2_mul(a[2], c1), fq2_mul(a[1], c2))), fq2_mul(a[0], c0)))
    return (fq2_mul(t, c0), fq2_mul(t, c1), fq2_mul(t, c2))


def fq6_scale(a, k):
    # src/fields/fq6.rs:67-73
    return tuple(fq2_mul(x, k) for x in a)


def fq6_frob(a, power: int):
    # src/fields/fq6.rs:75-81
    return (
        fq2_frob(a[0], power),
        fq2_mul(fq2_frob(a[1], power), frob_gamma(power, 2)),
        fq2_mul(fq2_frob(a[2], power), frob_gamma(power, 4)),
    )


# --------------------------------------------------------------------------
# Fq12 = Fq6[w]/(w^2 - v).  src/fields/fq12.rs
# --------------------------------------------------------------------------
FQ12_ONE = (FQ6_ONE, FQ6_ZERO)


def fq12_add(a, b):
    return (fq6_add(a[0], b[0]), fq6_add(a[1], b[1]))


def fq12_sub(a, b):
    return (fq6_sub(a[0], b[0]), fq6_sub(a[1], b[1]))


def fq12_neg(a):
    return (fq6_neg(a[0]), fq6_neg(a[1]))


def fq12_mul(a, b):
    # src/fields/fq12.rs:295-307
    aa = fq6_mul(a[0], b[0])
    bb = fq6_mul(a[1], b[1])
    return (
        fq6_add(fq6_mul_by_nonresidue(bb), aa),
        fq6_sub(fq6_sub(fq6_mul(fq6_add(a[0], a[1]), fq6_add(b[0], b[1])), aa), bb),
    )


def fq12_sqr(a):
    # src/fields/fq12.rs:275-282
    ab = fq6_mul(a[0], a[1])
    return (
        fq6_sub(
            fq6_sub(fq6_mul(fq6_add(fq6_mul_by_nonresidue(a[1]), a[0]), fq6_add(a[0], a[1])), ab),
            fq6_mul_by_nonresidue(ab)),
        fq6_add(ab, ab),
    )


def fq12_inv(a):
    # src/fields/fq12.rs:284-292
    t = fq6_inv(fq6_sub(fq6_sqr(a[0]), fq6_mul_by_nonresidue(fq6_sqr(a[1]))))
    return (fq6_mul(a[0], t), fq6_neg(fq6_mul(a[1], t)))


def fq12_conj(a):
    # unitary_inverse, src/fields/fq12.rs:103-105
    return (a[0], fq6_neg(a[1]))


def fq12_frob(a, power: int):
    # src/fields/fq12.rs:90-95
    return (fq6_frob(a[0], power), fq6_scale(fq6_frob(a[1], power), frob_gamma(power, 1)))


def fq12_mul_by_024(a, ell_0, ell_vw, ell_vv):
    # src/fields/fq12.rs:107-176 (literal statement order)
    z0, z1, z2 = a[0]
    z3, z4, z5 = a[1]
    x0, x2, x4 = ell_0, ell_vv, ell_vw
    d0 = fq2_mul(z0, x0)
    d2 = fq2_mul(z2, x2)
    d4 = fq2_mul(z4, x4)
    t2 = fq2_add(z0, z4)
    t1 = fq2_add(z0, z2)
    s0 = fq2_add(fq2_add(z1, z3), z5)
    s1 = fq2_mul(z1, x2)
    t3 = fq2_add(s1, d4)
    t4 = fq2_add(fq2_mul_xi(t3), d0)
    o0 = t4
    t3 = fq2_mul(z5, x4)
    s1 = fq2_add(s1, t3)
    t3 = fq2_add(t3, d2)
    t4 = fq2_mul_xi(t3)
    t3 = fq2_mul(z1, x0)
    s1 = fq2_add(s1, t3)
    t4 = fq2_add(t4, t3)
    o1 = t4
    t0 = fq2_add(x0, x2)
    t3 = fq2_sub(fq2_sub(fq2_mul(t1, t0), d0), d2)
    t4 = fq2_mul(z3, x4)
    s1 = fq2_add(s1, t4)
    t3 = fq2_add(t3, t4)
    t0 = fq2_add(z2, z4)
    o2 = t3
    t1 = fq2_add(x2, x4)
    t3 = fq2_sub(fq2_sub(fq2_mul(t0, t1), d2), d4)
    t4 = fq2_mul_xi(t3)
    t3 = fq2_mul(z3, x0)
    s1 = fq2_add(s1, t3)
    t4 = fq2_add(t4, t3)
    o3 = t4
    t3 = fq2_mul(z5, x2)
    s1 = fq2_add(s1, t3)
    t4 = fq2_mul_xi(t3)
    t0 = fq2_add(x0, x4)
    t3 = fq2_sub(fq2_sub(fq2_mul(t2, t0), d0), d4)
    t4 = fq2_add(t4, t3)
    o4 = t4
    t0 = fq2_add(fq2_add(x0, x2), x4)
    t3 = fq2_sub(fq2_mul(s0, t0), s1)
    o5 = t3
    return ((o0, o1, o2), (o3, o4, o5))


def fq12_cyclotomic_squared(a):
    # src/fields/fq12.rs:178-227 (literal Granger-Scott; pinned by test_cyclotomic_exp)
    z0, z4, z3 = a[0]
    z2, z1, z5 = a[1]

    def fp4_sqr(x, y):
        tmp = fq2_mul(x, y)
        t_re = fq2_sub(
            fq2_sub(fq2_mul(fq2_add(x, y), fq2_add(fq2_mul_xi(y), x)), tmp), fq2_mul_xi(tmp))
        return t_re, fq2_add(tmp, tmp)

    t0, t1 = fp4_sqr(z0, z1)
    t2, t3 = fp4_sqr(z2, z3)
    t4, t5 = fp4_sqr(z4, z5)

    def minus(t, z):  # 2*(t - z) + t
        d = fq2_sub(t, z)
        return fq2_add(fq2_add(d, d), t)

    def plus(t, z):  # 2*(t + z) + t
        s = fq2_add(t, z)
        return fq2_add(fq2_add(s, s), t)

    n0 = minus(t0, z0)
    n1 = plus(t1, z1)
    n2 = plus(fq2_mul_xi(t5), z2)
    n3 = minus(t4, z3)
    n4 = minus(t2, z4)
    n5 = plus(t3, z5)
    return ((n0, n4, n3), (n2, n1, n5))


def fq12_cyclotomic_pow(a, e: int):
    # src/fields/fq12.rs:229-246
    res = FQ12_ONE
    found_one = False
    for i in range(255, -1, -1):
        bit = (e >> i) & 1
        if found_one:
            res = fq12_cyclotomic_squared(res)
        if bit:
            found_one = True
            res = fq12_mul(a, res)
    return res


def fq12_exp_by_neg_z(a):
    # src/fields/fq12.rs:97-101
    return fq12_conj(fq12_cyclotomic_pow(a, U))


def fq12_pow(a, e: int):
    # FieldElement::pow, src/fields/mod.rs:35-46 (all 256 bits)
    res = FQ12_ONE
    for i in range(255, -1, -1):
        res = fq12_sqr(res)
        if (e >> i) & 1:
            res = fq12_mul(a, res)
    return res


def final_exp_first_chunk(f):
    # src/fields/fq12.rs:41-52
    b = fq12_inv(f)
    a = fq12_conj(f)
    c = fq12_mul(a, b)
    d = fq12_frob(c, 2)
    return fq12_mul(d, c)


def final_exp_last_chunk(s):
    # src/fields/fq12.rs:54-84
    a = fq12_exp_by_neg_z(s)
    b = fq12_cyclotomic_squared(a)
    c = fq12_cyclotomic_squared(b)
    d = fq12_mul(c, b)
    e = fq12_exp_by_neg_z(d)
    f = fq12_cyclotomic_squared(e)
    g = fq12_exp_by_neg_z(f)
    h = fq12_conj(d)
    i = fq12_conj(g)
    j = fq12_mul(i, e)
    k = fq12_mul(j, h)
    l = fq12_mul(k, b)
    m = fq12_mul(k, e)
    n = fq12_mul(s, m)
    o = fq12_frob(l, 1)
    p = fq12_mul(o, n)
    q = fq12_frob(k, 2)
    r = fq12_mul(q, p)
    ss = fq12_conj(s)
    t = fq12_mul(ss, l)
    u = fq12_frob(t, 3)
    return fq12_mul(u, r)


def final_exponentiation(f):
    # src/fields/fq12.rs:86-88
    return final_exp_last_chunk(final_exp_first_chunk(f))


# --------------------------------------------------------------------------
# Jacobian groups.  src/groups/mod.rs:83-140, 207-347
# A "field" is a small namespace of callables so G1 (Fq) and G2 (Fq2) share code.
# --------------------------------------------------------------------------
class _FqOps:
    zero = 0
    one = 1

    @staticmethod
    def add(a, b):
        return (a + b) % Q

    @staticmethod
    def sub(a, b):
        return (a - b) % Q

    @staticmethod
    def neg(a):
        return (-a) % Q

    @staticmethod
    def mul(a, b):
        return (a * b) % Q

    @staticmethod
    def sqr(a):
        return (a * a) % Q

    @staticmethod
    def inv(a):
        return fq_inv(a)

    @staticmethod
    def is_zero(a):
        return a % Q == 0


class _Fq2Ops:
    zero = FQ2_ZERO
    one = FQ2_ONE
    add = staticmethod(fq2_add)
    sub = staticmethod(fq2_sub)
    neg = staticmethod(fq2_neg)
    mul = staticmethod(fq2_mul)
    sqr = staticmethod(fq2_sqr)
    inv = staticmethod(fq2_inv)

    @staticmethod
    def is_zero(a):
        return a[0] % Q == 0 and a[1] % Q == 0


G1_GEN = (1, 2, 1)  # src/groups/mod.rs:356-362
G1_B = 3  # :364-366
# G2 generator, src/groups/mod.rs:378-390 (plain-domain values, standard alt_bn128 G2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
    FQ2_ONE,
)
G2_B = fq2_mul((3, 0), fq2_inv(XI))  # 3/xi, src/groups/mod.rs:392-397
TWO_INV = pow(2, -1, Q)  # src/groups/mod.rs:446-449


def g_zero(F):
    # src/groups/mod.rs:208-214
    return (F.zero, F.one, F.zero)


def g_is_zero(F, p):
    return F.is_zero(p[2])


def g_double(F, p):
    # src/groups/mod.rs:228-247
    x, y, z = p
    a = F.sqr(x)
    b = F.sqr(y)
    c = F.sqr(b)
    d = F.sub(F.sub(F.sqr(F.add(x, b)), a), c)
    d = F.add(d, d)
    e = F.add(F.add(a, a), a)
    f = F.sqr(e)
    x3 = F.sub(f, F.add(d, d))
    eight_c = F.add(c, c)
    eight_c = F.add(eight_c, eight_c)
    eight_c = F.add(eight_c, eight_c)
    y1z1 = F.mul(y, z)
    return (x3, F.sub(F.mul(e, F.sub(d, x3)), eight_c), F.add(y1z1, y1z1))


def g_add(F, p, o):
    # src/groups/mod.rs:272-312
    if g_is_zero(F, p):
        return o
    if g_is_zero(F, o):
        return p
    z1_squared = F.sqr(p[2])
    z2_squared = F.sqr(o[2])
    u1 = F.mul(p[0], z2_squared)
    u2 = F.mul(o[0], z1_squared)
    z1_cubed = F.mul(p[2], z1_squared)
    z2_cubed = F.mul(o[2], z2_squared)
    s1 = F.mul(p[1], z2_cubed)
    s2 = F.mul(o[1], z1_cubed)
    if u1 == u2 and s1 == s2:
        return g_double(F, p)
    h = F.sub(u2, u1)
    s2_minus_s1 = F.sub(s2, s1)
    i = F.sqr(F.add(h, h))
    j = F.mul(h, i)
    r = F.add(s2_minus_s1, s2_minus_s1)
    v = F.mul(u1, i)
    s1_j = F.mul(s1, j)
    x3 = F.sub(F.sub(F.sqr(r), j), F.add(v, v))
    return (
        x3,
        F.sub(F.mul(r, F.sub(v, x3)), F.add(s1_j, s1_j)),
        F.mul(F.sub(F.sub(F.sqr(F.add(p[2], o[2])), z1_squared), z2_squared), h),
    )


def g_neg(F, p):
    # src/groups/mod.rs:314-328
    if g_is_zero(F, p):
        return p
    return (p[0], F.neg(p[1]), p[2])


def g_mul(F, p, k: int):
    """`G * Fr` with k the de-Montgomerized scalar (src/groups/mod.rs:250-270)."""
    res = g_zero(F)
    found_one = False
    for i in range(255, -1, -1):
        if found_one:
            res = g_double(F, res)
        if (k >> i) & 1:
            found_one = True
            res = g_add(F, res, p)
    return res


def g_eq(F, p, o):
    # src/groups/mod.rs:83-109
    if g_is_zero(F, p):
        return g_is_zero(F, o)
    if g_is_zero(F, o):
        return False
    z1s, z2s = F.sqr(p[2]), F.sqr(o[2])
    if F.mul(p[0], z2s) != F.mul(o[0], z1s):
        return False
    return F.mul(p[1], F.mul(o[2], z2s)) == F.mul(o[1], F.mul(p[2], z1s))


def g_to_affine(F, p):
    # src/groups/mod.rs:113-130
    if F.is_zero(p[2]):
        return None
    if p[2] == F.one:
        return (p[0], p[1])
    zinv = F.inv(p[2])
    zinv2 = F.sqr(zinv)
    return (F.mul(p[0], zinv2), F.mul(p[1], F.mul(zinv2, zinv)))


def g_normalize(F, p):
    # Group::normalize, src/lib.rs:88-95
    a = g_to_affine(F, p)
    return p if a is None else (a[0], a[1], F.one)


FQ = _FqOps
FQ2 = _Fq2Ops

# --------------------------------------------------------------------------
# Optimal ate pairing.  src/groups/mod.rs:441-634, 764-771
# --------------------------------------------------------------------------


def _ate_bits():
    bits = bin(ATE_LOOP_COUNT)[2:]
    return [int(b) for b in bits[1:]]  # MSB skipped, src/groups/mod.rs:493-498


def g2_mul_by_q(a):
    # src/groups/mod.rs:550-555 ; constants :456-470 = xi^((q-1)/3), xi^((q-1)/2)
    return (fq2_mul(frob_gamma(1, 2), fq2_frob(a[0], 1)), fq2_mul(frob_gamma(1, 3), fq2_frob(a[1], 1)))


def _doubling_step(r):
    # src/groups/mod.rs:612-634 ; returns (new_r, (ell_0, ell_vw, ell_vv))
    x, y, z = r
    a = fq2_scale(fq2_mul(x, y), TWO_INV)
    b = fq2_sqr(y)
    c = fq2_sqr(z)
    d = fq2_add(fq2_add(c, c), c)
    e = fq2_mul(G2_B, d)
    f = fq2_add(fq2_add(e, e), e)
    g = fq2_scale(fq2_add(b, f), TWO_INV)
    h = fq2_sub(fq2_sqr(fq2_add(y, z)), fq2_add(b, c))
    i = fq2_sub(e, b)
    j = fq2_sqr(x)
    e_sq = fq2_sqr(e)
    nx = fq2_mul(a, fq2_sub(b, f))
    ny = fq2_sub(fq2_sqr(g), fq2_add(fq2_add(e_sq, e_sq), e_sq))
    nz = fq2_mul(b, h)
    return (nx, ny, nz), (fq2_mul(XI, i), fq2_neg(h), fq2_add(fq2_add(j, j), j))


def _mixed_addition_step(r, base):
    # src/groups/mod.rs:592-610
    x, y, z = r
    d = fq2_sub(x, fq2_mul(z, base[0]))
    e = fq2_sub(y, fq2_mul(z, base[1]))
    f = fq2_sqr(d)
    g = fq2_sqr(e)
    h = fq2_mul(d, f)
    i = fq2_mul(x, f)
    j = fq2_sub(fq2_add(fq2_mul(z, g), h), fq2_add(i, i))
    nx = fq2_mul(d, j)
    ny = fq2_sub(fq2_mul(e, fq2_sub(i, j)), fq2_mul(h, y))
    nz = fq2_mul(z, h)
    ell_0 = fq2_mul(XI, fq2_sub(fq2_mul(e, base[0]), fq2_mul(d, base[1])))
    return (nx, ny, nz), (ell_0, d, fq2_neg(e))  # (ell_0, ell_vw, ell_vv)


def g2_precompute(q_aff):
    """AffineG<G2>::precompute, src/groups/mod.rs:557-588 -> list of 102 (ell_0, ell_vw, ell_vv)."""
    r = (q_aff[0], q_aff[1], FQ2_ONE)
    coeffs = []
    for bit in _ate_bits():
        r, c = _doubling_step(r)
        coeffs.append(c)
        if bit:
            r, c = _mixed_addition_step(r, q_aff)
            coeffs.append(c)
    q1 = g2_mul_by_q(q_aff)
    q2m = g2_mul_by_q(q1)
    q2 = (q2m[0], fq2_neg(q2m[1]))
    r, c = _mixed_addition_step(r, q1)
    coeffs.append(c)
    r, c = _mixed_addition_step(r, q2)
    coeffs.append(c)
    return coeffs


def miller_loop(coeffs, p_aff):
    # G2Precomp::miller_loop, src/groups/mod.rs:485-520
    f = FQ12_ONE
    idx = 0
    px, py = p_aff
    for bit in _ate_bits():
        c = coeffs[idx]
        idx += 1
        f = fq12_mul_by_024(fq12_sqr(f), c[0], fq2_scale(c[1], py), fq2_scale(c[2], px))
        if bit:
            c = coeffs[idx]
            idx += 1
            f = fq12_mul_by_024(f, c[0], fq2_scale(c[1], py), fq2_scale(c[2], px))
    for _ in range(2):
        c = coeffs[idx]
        idx += 1
        f = fq12_mul_by_024(f, c[0], fq2_scale(c[1], py), fq2_scale(c[2], px))
    assert idx == len(coeffs) == 102
    return f


def pairing(p, q):
    """groups::pairing, src/groups/mod.rs:764-771.  p: Jacobian G1 triple, q: Jacobian G2 triple."""
    pa = g_to_affine(FQ, p)
    qa = g_to_affine(FQ2, q)
    if pa is None or qa is None:
        return FQ12_ONE
    return final_exponentiation(miller_loop(g2_precompute(qa), pa))


# --------------------------------------------------------------------------
# Byte images == the crate's #[repr(C)] layouts:
# Montgomery form, canonical, 4 little-endian u64 limbs per Fq/Fr.
# --------------------------------------------------------------------------
def _limbs_bytes(x: int) -> bytes:
    return x.to_bytes(32, "little")


def fq_to_bytes(x: int, p: int = Q) -> bytes:
    return _limbs_bytes(to_mont(x % p, p))


def fq_from_bytes(b: bytes, p: int = Q) -> int:
    return from_mont(int.from_bytes(b[:32], "little"), p)


def fr_to_bytes(x: int) -> bytes:
    return fq_to_bytes(x, R_ORDER)


def fr_from_bytes(b: bytes) -> int:
    return fq_from_bytes(b, R_ORDER)


def fq2_to_bytes(a) -> bytes:
    return fq_to_bytes(a[0]) + fq_to_bytes(a[1])


def fq2_from_bytes(b: bytes):
    return (fq_from_bytes(b[0:32]), fq_from_bytes(b[32:64]))


def g1_to_bytes(p) -> bytes:
    return b"".join(fq_to_bytes(c) for c in p)


def g1_from_bytes(b: bytes):
    return tuple(fq_from_bytes(b[32 * i:32 * i + 32]) for i in range(3))


def g2_to_bytes(p) -> bytes:
    return b"".join(fq2_to_bytes(c) for c in p)


def g2_from_bytes(b: bytes):
    return tuple(fq2_from_bytes(b[64 * i:64 * i + 64]) for i in range(3))


def gt_to_bytes(f) -> bytes:
    return b"".join(fq2_to_bytes(f[h][i]) for h in range(2) for i in range(3))


def gt_from_bytes(b: bytes):
    cs = [fq2_from_bytes(b[64 * k:64 * k + 64]) for k in range(6)]
    return ((cs[0], cs[1], cs[2]), (cs[3], cs[4], cs[5]))


def fq12_flat(f):
    """12 plain integers in reference declaration order (c0.c0.c0, c0.c0.c1, c0.c1.c0, ...)."""
    return [f[h][i][j] for h in range(2) for i in range(3) for j in range(2)]


def fq12_from_flat(v):
    v = [int(x) for x in v]
    return (((v[0], v[1]), (v[2], v[3]), (v[4], v[5])), ((v[6], v[7]), (v[8], v[9]), (v[10], v[11])))


# --------------------------------------------------------------------------
# Wire format (host side).  src/arith.rs:100-159, src/fields/fq2.rs:31-53,
# src/groups/mod.rs:143-176
# --------------------------------------------------------------------------
def encode_fr(x: int) -> bytes:
    return (x % R_ORDER).to_bytes(32, "big")


def encode_g1(p) -> bytes:
    a = g_to_affine(FQ, p)
    if a is None:
        return b"\x00"
    return b"\x04" + a[0].to_bytes(32, "big") + a[1].to_bytes(32, "big")


def encode_g2(p) -> bytes:
    a = g_to_affine(FQ2, p)
    if a is None:
        return b"\x00"

    def enc2(c):
        return (c[1] * Q + c[0]).to_bytes(64, "big")

    return b"\x04" + enc2(a[0]) + enc2(a[1])


def decode_g1(b: bytes):
    if b[0] == 0:
        return g_zero(FQ)
    if b[0] != 4:
        raise ValueError("invalid leading byte for uncompressed group element")
    x = int.from_bytes(b[1:33], "big")
    y = int.from_bytes(b[33:65], "big")
    if x >= Q or y >= Q:
        raise ValueError("integer is not less than modulus")
    if (y * y - (x * x * x + G1_B)) % Q != 0:
        raise ValueError("point is not on the curve")
    return (x, y, 1)


def decode_g2(b: bytes):
    if b[0] == 0:
        return g_zero(FQ2)
    if b[0] != 4:
        raise ValueError("invalid leading byte for uncompressed group element")

    def dec2(bb):
        c1, c0 = divmod(int.from_bytes(bb, "big"), Q)
        if c1 >= Q:
            raise ValueError("integer not less than modulus squared")
        return (c0, c1)

    x = dec2(b[1:65])
    y = dec2(b[65:129])
    if fq2_sqr(y) != fq2_add(fq2_mul(fq2_sqr(x), x), G2_B):
        raise ValueError("point is not on the curve")
    p = (x, y, FQ2_ONE)
    if not g_is_zero(FQ2, g_add(FQ2, g_mul(FQ2, p, R_ORDER - 1), p)):
        raise ValueError("point is not in the subgroup")
    return p


# --------------------------------------------------------------------------
# Deterministic synthetic inputs: splitmix64 -> 512 bits -> mod r.
# Shared by the C oracle (oracle/bn_ref.c), the tests and bench.py.
# --------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def splitmix64(state: int):
    state = (state + 0x9E3779B97F4A7C15) & _M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return state, z ^ (z >> 31)


def synth_scalar(seed: int, index: int, p: int = R_ORDER) -> int:
    """Uniform-ish residue: 8 splitmix64 words (little-endian limbs) mod p (cf. src/arith.rs:195-198)."""
    state = (seed * 0x9E3779B97F4A7C15 + index * 0xD1B54A32D192ED03) & _M64
    v = 0
    for k in range(8):
        state, w = splitmix64(state)
        v |= w << (64 * k)
    return v % p
