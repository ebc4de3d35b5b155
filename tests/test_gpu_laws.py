"""The reference's algebraic-law suites re-expressed as batches on the GPU (own seeds), through the C ABI:
field_trials (src/fields/tests.rs:4-128) on Fr (k_fr_op / k_fr_pow) and on Gt = Fq12 (k_gt_mul / k_gt_inv / k_gt_pow),
group_trials (src/groups/tests.rs:5-102) on G1 and G2 (k_g{1,2}_op, k_g{1,2}_mul, k_g{1,2}_eq), and the golden
recurrences of tests/serialization.rs replayed ENTIRELY on the device (no oracle arithmetic inside the loops)."""
import numpy as np
import pytest

from oracle import bn_oracle as o
from oracle import cref
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bn():
    import bn_b200
    bn_b200.init(0)
    return bn_b200


def fr(bn, op, a, b=None):
    return bn.fr_op_batch(op, a, b)


def rep(img, n):
    return np.repeat(np.asarray(img, dtype=np.uint64)[None], n, axis=0)


# ---------------------------------------------------------------------------------------------------------------------
# field_trials::<Fr>
# ---------------------------------------------------------------------------------------------------------------------
def test_fr_field_trials(bn):
    one, zero = util.fr_img(1), util.fr_img(0)
    # can_invert: a = 1, 2, ..., 10000 and a = -1, -2, ...: a * a^-1 == 1; zero has no inverse (device returns 0)
    n = 10000
    up = np.stack([util.fr_img(i) for i in range(1, n + 1)])
    down = np.stack([util.fr_img(o.R_ORDER - i) for i in range(1, n + 1)])
    for a in (up, down):
        assert (fr(bn, "mul", a, fr(bn, "inverse", a)) == one[None]).all()
    assert not fr(bn, "inverse", zero[None]).any()
    assert not fr(bn, "neg", zero[None]).any()                                   # -0 == 0
    assert not fr(bn, "add", fr(bn, "neg", one[None]), one[None]).any()          # -1 + 1 == 0
    assert not fr(bn, "sub", zero[None], zero[None]).any()
    m = 1000
    a, b, c, d = (util.synth_scalars(0xF1E1D0 + i, m) for i in range(4))
    # rand_element_squaring: a*a == a.squared() == a.pow(2); counting up from zero too
    two = rep(util.fr_img(2), m)
    assert np.array_equal(fr(bn, "mul", a, a), bn.fr_pow_batch(a, two))
    cnt = np.stack([util.fr_img(i) for i in range(100)])
    assert np.array_equal(fr(bn, "mul", cnt, cnt), bn.fr_pow_batch(cnt, two[:100]))
    # rand_element_addition_and_negation
    assert not fr(bn, "add", a, fr(bn, "neg", a)).any()
    x, r0 = a.copy(), b.copy()
    y = fr(bn, "add", x, r0)
    for k in range(10):
        r = [util.synth_scalars(0xADD0 + 16 * k + j, m) for j in range(5)]
        x, y = fr(bn, "add", x, r[0]), fr(bn, "add", y, r[0])
        x, y = fr(bn, "sub", x, r[1]), fr(bn, "sub", y, r[1])
        nn = fr(bn, "neg", fr(bn, "neg", r[2]))
        x, y = fr(bn, "add", x, nn), fr(bn, "add", y, nn)
        x, y = fr(bn, "sub", x, r[3]), fr(bn, "add", y, fr(bn, "neg", r[3]))
        x, y = fr(bn, "add", x, fr(bn, "neg", r[4])), fr(bn, "sub", y, r[4])
    assert np.array_equal(x, fr(bn, "sub", y, r0))
    # rand_element_multiplication (associativity), rand_element_inverse, rand_element_eval (distributivity)
    assert np.array_equal(fr(bn, "mul", fr(bn, "mul", a, b), c), fr(bn, "mul", a, fr(bn, "mul", b, c)))
    ai = fr(bn, "inverse", a)
    assert (fr(bn, "mul", ai, a) == one[None]).all()
    assert np.array_equal(fr(bn, "mul", fr(bn, "mul", a, b), ai), b)
    lhs = fr(bn, "mul", fr(bn, "add", a, b), fr(bn, "add", c, d))
    rhs = fr(bn, "add", fr(bn, "add", fr(bn, "mul", a, c), fr(bn, "mul", b, c)), fr(bn, "add", fr(bn, "mul", a, d), fr(bn, "mul", b, d)))
    assert np.array_equal(lhs, rhs)
    # every op against the oracle on the same inputs, extremes included
    a[0], a[1], b[0], b[2] = zero, util.fr_img(o.R_ORDER - 1), util.fr_img(o.R_ORDER - 1), zero
    s = slice(0, 200)
    for op in ("mul", "add", "sub"):
        assert np.array_equal(fr(bn, op, a[s], b[s]), cref.fp_op(op, 1, a[s], b[s])), op
    assert np.array_equal(fr(bn, "neg", a[s]), cref.fp_op("neg", 1, a[s]))
    # Fr::pow against plain-integer pow (exponent = U256::from(e)); exponents 0, 1, r-1 included
    e = util.synth_scalars(0xE0, 64)
    e[0], e[1], e[2] = zero, one, util.fr_img(o.R_ORDER - 1)
    rinv = pow(1 << 256, -1, o.R_ORDER)
    plain = lambda w: int.from_bytes(w.tobytes(), "little") * rinv % o.R_ORDER
    want = np.stack([util.fr_img(pow(plain(x), plain(k), o.R_ORDER)) for x, k in zip(a[:64], e)])
    assert np.array_equal(bn.fr_pow_batch(a[:64], e), want)


# ---------------------------------------------------------------------------------------------------------------------
# field_trials on Gt = Fq12 (mul, inverse, pow are the crate's Gt API; src/lib.rs:165-179)
# ---------------------------------------------------------------------------------------------------------------------
def rand_fq12(seed, n):
    import random
    rng = random.Random(seed)
    return np.stack([util.gt_img(o.fq12_from_flat([rng.randrange(o.Q) for _ in range(12)])) for _ in range(n)])


def test_gt_field_trials(bn):
    n = 250
    a, b, c = rand_fq12(1, n), rand_fq12(2, n), rand_fq12(3, n)
    one = util.gt_img(o.FQ12_ONE)
    mul, inv = bn.gt_mul_batch, bn.gt_inv_batch
    assert np.array_equal(mul(mul(a, b), c), mul(a, mul(b, c)))                  # rand_element_multiplication
    assert np.array_equal(mul(a, b), mul(b, a))
    ai = inv(a)
    assert (mul(ai, a) == one[None]).all()                                       # rand_element_inverse
    assert np.array_equal(mul(mul(a, b), ai), b)
    assert np.array_equal(mul(a, rep(one, n)), a)
    two = rep(util.fr_img(2), n)
    assert np.array_equal(bn.gt_pow_batch(a, two), mul(a, a))                    # rand_element_squaring
    # distributivity needs Fq12 addition, which is not in the crate's Gt API: the sums come from the oracle, both
    # products from the GPU:  (a + b) * c == a*c + b*c
    lhs = mul(cref.fq12_add(a[:50], b[:50]), c[:50])
    assert np.array_equal(lhs, cref.fq12_add(mul(a[:50], c[:50]), mul(b[:50], c[:50])))
    # and everything against the oracle
    assert np.array_equal(mul(a[:60], b[:60]), cref.fq12_mul(a[:60], b[:60]))
    assert np.array_equal(ai[:60], cref.fq12_inv(a[:60]))


def test_cyclotomic_exp_kat_on_device(bn):
    """The reference's test_cyclotomic_exp (src/fields/mod.rs:171-201): exp_by_neg_z of a NON-cyclotomic element, which
    only the literal Granger-Scott formula reproduces -- through a device entry point (PTX carry chains, not the host
    emulator's C branches)."""
    kat = util.load_json("fq12_kat.json")
    x = util.gt_img(o.fq12_from_flat(kat["cyclotomic_orig"]))[None]
    want = util.gt_img(o.fq12_from_flat(kat["cyclotomic_expected"]))[None]
    assert np.array_equal(bn.gt_exp_by_neg_z_batch(x), want)
    # on pairing values (cyclotomic subgroup) it is f^-u: compare with the oracle and with Gt::pow + inverse
    g1, g2 = util.synth_pairs(0xC1C, 11)
    f = bn.pairing_batch(g1, g2)
    got = bn.gt_exp_by_neg_z_batch(f)
    assert np.array_equal(got, cref.fq12_exp_by_neg_z(f))
    u = rep(util.fr_img(4965661367192848881), len(f))
    assert np.array_equal(got, bn.gt_inv_batch(bn.gt_pow_batch(f, u)))


# ---------------------------------------------------------------------------------------------------------------------
# group_trials::<G1>, ::<G2>
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["g1", "g2"])
def test_group_trials(bn, kind):
    n = 50
    if kind == "g1":
        gen, zero, words = cref.g1_generator()[0], util.g1_img(o.g_zero(o.FQ)), 12
        gop, geq, gmul, norm = bn.g1_op_batch, bn.g1_eq_batch, bn.g1_mul_batch, bn.g1_normalize_batch
        c_add, c_neg, c_dbl = cref.g1_add, cref.g1_neg, cref.g1_double
    else:
        gen, zero, words = cref.g2_generator()[0], util.g2_img(o.g_zero(o.FQ2)), 24
        gop, geq, gmul, norm = bn.g2_op_batch, bn.g2_eq_batch, bn.g2_mul_batch, bn.g2_normalize_batch
        c_add, c_neg, c_dbl = cref.g2_add, cref.g2_neg, cref.g2_double
    zoff = 2 * words // 3
    is_zero = lambda p: ~p[:, zoff:].any(axis=1)
    add = lambda a, b: gop("add", a, b)
    sub = lambda a, b: gop("sub", a, b)
    neg = lambda a: gop("neg", a)
    dbl = lambda a: gop("double", a)
    rnd = lambda seed: gmul(rep(gen, n), util.synth_scalars(seed, n))           # G::random = one * Fr::random
    one, frone = rep(gen, 1), util.fr_img(1)[None]
    assert is_zero(zero[None]).all()
    assert is_zero(sub(one, one)).all()
    assert geq(add(one, one), gmul(one, util.fr_img(2)[None])).all()
    assert is_zero(dbl(zero[None])).all()
    assert is_zero(add(gmul(one, fr(bn, "neg", frone)), one)).all()
    r1, r2, r3 = rnd(0x6101), rnd(0x6102), rnd(0x6103)
    # random_test_addition
    assert geq(add(add(r1, r2), r3), add(r1, add(r2, r3))).all()
    assert is_zero(sub(sub(sub(add(add(r1, r2), r3), r2), r3), r1)).all()
    # random_test_doubling
    ti = fr(bn, "inverse", rep(util.fr_img(2), n))
    assert geq(add(add(r1, r2), r1), add(dbl(r1), r2)).all()
    assert geq(r1, gmul(dbl(r1), ti)).all()
    # random_test_dh
    ska, skb = util.synth_scalars(0x6104, n), util.synth_scalars(0x6105, n)
    assert geq(gmul(gmul(rep(gen, n), skb), ska), gmul(gmul(rep(gen, n), ska), skb)).all()
    # random_test_equality
    begin, a, b, c, d = rnd(0x6106), util.synth_scalars(0x6107, n), rnd(0x6108), util.synth_scalars(0x6109, n), rnd(0x610A)
    acc = begin
    for _ in range(10):
        acc = dbl(sub(neg(gmul(add(neg(gmul(acc, a)), b), c)), d))
    ai, ci = fr(bn, "inverse", a), fr(bn, "inverse", c)
    for _ in range(10):
        acc = gmul(neg(sub(gmul(neg(add(gmul(acc, ti), d)), ci), b)), ai)
    assert geq(acc, begin).all()
    assert not geq(acc, r1).any()
    # limb-exact against the oracle, corner cases of src/groups/mod.rs:272-347 included:
    #   zero + p, p + zero, p + p (-> double), p + (-p) (non-canonical z = 0 triple), neg(zero), double(zero)
    x = np.concatenate([r1[:20], zero[None], r1[:1], r1[1:2], r1[2:3], zero[None], norm(r1[3:4])])
    y = np.concatenate([r2[:20], r1[:1], zero[None], r1[1:2], neg(r1[2:3]), zero[None], r1[3:4]])
    assert np.array_equal(add(x, y), c_add(x, y))
    assert np.array_equal(sub(x, y), c_add(x, c_neg(y)))
    assert np.array_equal(neg(x), c_neg(x))
    assert np.array_equal(dbl(x), c_dbl(x))
    assert geq(x[-1:], y[-1:]).all() and geq(zero[None], sub(r1[:1], r1[:1])).all() and not geq(zero[None], r1[:1]).any()
    # the value classes mirror the crate's operators
    G = bn.G1 if kind == "g1" else bn.G2
    p, q = G(r1[0]), G(r2[0])
    assert (p + q) - q == p and (-p) + p == G.zero() and G.zero().is_zero() and not G.one().is_zero()
    assert G.one() * bn.Fr.from_int(2) == G.one() + G.one() and p.double() == p + p
    t = G(r1[0]); t.normalize()
    assert t == p and np.array_equal(t.img[zoff:], G.one().img[zoff:])
    s1, s2 = bn.Fr.from_int(12345), bn.Fr.from_int(67890)
    assert (s1 * s2) * s1.inverse() == s2 and s1 + (-s1) == bn.Fr.zero() and s1 - s1 == bn.Fr.zero()
    assert s1.pow(bn.Fr.from_int(3)) == s1 * s1 * s1 and bn.Fr.zero().inverse() is None and bn.Fr.one() * s1 == s1


# ---------------------------------------------------------------------------------------------------------------------
# golden recurrences of tests/serialization.rs, every operation on the device
# ---------------------------------------------------------------------------------------------------------------------
def test_g1_golden_recurrence_all_on_device(bn):
    """g1_vectors (tests/serialization.rs:74-10087): acc <- acc * 23938123 + acc, 2000 committed steps + the 10000th
    element; scalar-mul, addition, normalisation and encoding all run on the GPU."""
    lines = util.load_vectors("g1_vectors.txt")
    k = util.fr_img(23938123)[None]
    acc, accs = cref.g1_generator(), []
    last = util.load_json("last_vectors.json")["g1_9999"]
    for i in range(10000):
        if i < len(lines):
            accs.append(acc[0])
        if i == 9999:
            accs.append(acc[0])
            break
        acc = bn.g1_op_batch("add", bn.g1_mul_batch(acc, k), acc)
    enc = bn.encode_batch("g1", np.stack(accs))
    for i, want in enumerate(lines):
        assert bn.to_wire("g1", enc[i]).hex() == want, i
    assert bn.to_wire("g1", enc[-1]).hex() == last


def test_g2_golden_recurrence_all_on_device(bn):
    lines = util.load_vectors("g2_vectors.txt")
    k = util.fr_img(23938123)[None]
    acc, accs = cref.g2_generator(), []
    last = util.load_json("last_vectors.json")["g2_9999"]
    for i in range(10000):
        if i < len(lines):
            accs.append(acc[0])
        if i == 9999:
            accs.append(acc[0])
            break
        acc = bn.g2_op_batch("add", bn.g2_mul_batch(acc, k), acc)
    enc = bn.encode_batch("g2", np.stack(accs))
    for i, want in enumerate(lines):
        assert bn.to_wire("g2", enc[i]).hex() == want, i
    assert bn.to_wire("g2", enc[-1]).hex() == last


def test_fr_golden_recurrence_all_on_device(bn):
    """fr_vectors (tests/serialization.rs:20131-30143): acc <- acc*acc + acc + acc^-1, all 10000 steps on the GPU."""
    lines = util.load_vectors("fr_vectors.txt")
    acc, accs = util.fr_img(1)[None], []
    for _ in lines:
        accs.append(acc[0])
        sq = bn.fr_op_batch("mul", acc, acc)
        acc = bn.fr_op_batch("add", bn.fr_op_batch("add", sq, acc), bn.fr_op_batch("inverse", acc))
    enc = bn.encode_batch("fr", np.stack(accs))
    for i, want in enumerate(lines):
        assert bytes(enc[i]).hex() == want, i
