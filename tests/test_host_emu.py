"""CPU tests: the kernels' arithmetic (bn_b200/csrc/*.cuh) compiled for the host -- PTX chains swapped for
their portable equivalents, warp shuffles for a six-thread barrier exchange -- against the oracle.
This validates the tower formulas, the hexad lane choreography and the line schedule without a GPU;
the -m gpu tests validate the real kernels through the C ABI."""
import random

import numpy as np
import pytest

from oracle import bn_oracle as o
from oracle import cref
from tests import emu, util

rnd = random.Random(20260925)


def _int(a):
    return int.from_bytes(np.asarray(a, dtype="<u8").tobytes(), "little")


def _w(x):
    return util.words(x.to_bytes(32, "little"))


def fq2img(a):
    return util.words(o.fq2_to_bytes(a))


@pytest.mark.parametrize("which,p", [(0, o.Q), (1, o.R_ORDER)])
def test_fp_ops(which, p):
    rinv = pow(o.MONT_R, -1, p)
    specials = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 2**253, 2**32 - 1, (2**256 - 1) % p]
    vals = specials + [rnd.randrange(p) for _ in range(150)]
    for a in vals:
        b = rnd.choice(vals)
        assert _int(emu.fp_op(0, which, _w(a), _w(b))) == a * b * rinv % p
        assert _int(emu.fp_op(1, which, _w(a), _w(b))) == (a + b) % p
        assert _int(emu.fp_op(2, which, _w(a), _w(b))) == (a - b) % p
        assert _int(emu.fp_op(3, which, _w(a))) == (-a) % p
        assert _int(emu.fp_op(5, which, _w(a))) == a * pow(2, -1, p) % p
        assert _int(emu.fp_op(6, which, _w(a))) == a * rinv % p
    for a in [1, 2, p - 1] + [rnd.randrange(1, p) for _ in range(6)]:
        assert _int(emu.fp_op(4, which, _w(a))) == pow(a * rinv % p, -1, p) * o.MONT_R % p
    if which == 0:  # binary-Euclid inversion used by the single-thread batched-inversion step
        for a in [1, 2, 3, p - 1, p - 2, 2**253, (p + 1) // 2] + [rnd.randrange(1, p) for _ in range(200)]:
            assert _int(emu.fp_op(7, 0, _w(a))) == pow(a * rinv % p, -1, p) * o.MONT_R % p


def test_interleaved_multiplication_at_its_operand_bounds():
    """fp.cuh mul_reduce_rows (product and Montgomery reduction rows on one accumulator): canonical results for the lazy
    operands its callers pass -- q itself (the lazy negation of 0), 2q - 1 (the sums of duo.cuh's squaring) -- and for
    a*b + c*d at its bound 2 q^2."""
    q = o.Q
    rinv = pow(o.MONT_R, -1, q)
    big = [q, q - 1, 1, 0, 2**253, (q + 1) // 2] + [rnd.randrange(q) for _ in range(40)]
    for a in big:
        for b in (q, q - 1, rnd.choice(big)):
            c, d = rnd.choice(big), rnd.choice(big)
            assert _int(emu.fp_mul2(_w(a), _w(b), _w(c), _w(d))) == (a * b + c * d) * rinv % q
    assert _int(emu.fp_mul2(_w(q), _w(q), _w(q), _w(q))) == 0
    for x in (2 * q - 1, 2 * q - 2, q + 1, rnd.randrange(q, 2 * q)):
        for y in (2 * q - 1, q, rnd.randrange(2 * q)):
            assert _int(emu.fp_op(0, 0, _w(x), _w(y))) == x * y * rinv % q


def test_fp2_ops():
    edge = [(0, 0), (o.Q - 1, o.Q - 1), (5, 0), (0, o.Q - 1), (1, 1)]
    # values straddling every multiple of q in 9x - y / 9y + x, to stress the quotient estimate
    for kq in range(1, 10):
        for dlt in (-1, 0, 1):
            x = (kq * o.Q + dlt) * pow(9, -1, o.Q) % o.Q
            edge.append((x, 0))
            edge.append((0, x))
    for i in range(len(edge)):
        a = edge[i]
        assert np.array_equal(emu.fp2_op(6, fq2img(a)), fq2img(o.fq2_mul_xi(a))), a
    for i in range(120):
        a = edge[i % 5] if i < 10 else (rnd.randrange(o.Q), rnd.randrange(o.Q))
        b = edge[(i // 2) % 5] if i < 10 else (rnd.randrange(o.Q), rnd.randrange(o.Q))
        assert np.array_equal(emu.fp2_op(0, fq2img(a), fq2img(b)), fq2img(o.fq2_mul(a, b)))
        assert np.array_equal(emu.fp2_op(1, fq2img(a)), fq2img(o.fq2_sqr(a)))
        assert np.array_equal(emu.fp2_op(2, fq2img(a)), fq2img(o.fq2_mul_xi(a)))
        assert np.array_equal(emu.fp2_op(6, fq2img(a)), fq2img(o.fq2_mul_xi(a)))  # quotient-estimate reduction
        if a != (0, 0):
            assert np.array_equal(emu.fp2_op(3, fq2img(a)), fq2img(o.fq2_inv(a)))


@pytest.fixture(scope="module")
def pairs():
    g1, g2 = util.synth_pairs(3, 4)
    return g1, g2, cref.pairing_batch(g1, g2, 4)


def test_scalar_mul_limb_exact(pairs):
    g1, g2, _ = pairs
    ks = [o.synth_scalar(4, 0), 0, 1, o.R_ORDER - 1]
    for i, k in enumerate(ks):
        fr = util.fr_img(k)
        assert np.array_equal(emu.g1_mul(g1[i], fr), cref.g1_mul_batch(g1[i:i + 1], fr[None])[0])
        assert np.array_equal(emu.g2_mul(g2[i], fr), cref.g2_mul_batch(g2[i:i + 1], fr[None])[0])


def test_line_schedule_matches_precompute(pairs):
    """With the reference's binary walk (BN_ATE_NAF=0) the kernels' line code reproduces precompute() and the
    unreduced Miller value exactly (test_prepared_g2 / test_miller_loop granularity)."""
    g1, g2, _ = pairs
    assert emu.lib("refchain").emu_ate_naf() == 0 and emu.lib("refchain").emu_num_lines() == 102
    finite, L, pa, qa = emu.lines(g1[0], g2[0], "refchain")
    assert finite == 1
    P = o.g_to_affine(o.FQ, util.img_g1(g1[0]))
    Qa = o.g_to_affine(o.FQ2, util.img_g2(g2[0]))
    assert np.array_equal(pa, np.concatenate([util.words(o.fq_to_bytes(P[0])), util.words(o.fq_to_bytes(P[1]))]))
    assert np.array_equal(qa, np.concatenate([fq2img(Qa[0]), fq2img(Qa[1])]))
    co = o.g2_precompute(Qa)
    for t, (e0, evw, evv) in enumerate(co):
        l3 = o.fq2_scale(evw, P[1])
        l4 = o.fq2_scale(evv, P[0])
        exp = np.concatenate([fq2img(e0), fq2img(l3), fq2img(o.fq2_mul_xi(l3)), fq2img(l4), fq2img(o.fq2_mul_xi(l4))])
        assert np.array_equal(L[t], exp), t
    assert np.array_equal(emu.miller(L, "refchain"), util.gt_img(o.miller_loop(co, P)))
    assert np.array_equal(emu.pairing(g1[0], g2[0], "refchain"), pairs[2][0])


def test_naf_schedule_same_gt(pairs):
    """The shipped NAF walk (88 lines) changes the unreduced Miller value but not Gt."""
    g1, g2, gt = pairs
    assert emu.lib().emu_ate_naf() == 1 and emu.lib().emu_num_lines() == 88
    _, L, _, _ = emu.lines(g1[0], g2[0])
    P = o.g_to_affine(o.FQ, util.img_g1(g1[0]))
    Qa = o.g_to_affine(o.FQ2, util.img_g2(g2[0]))
    unreduced = emu.miller(L)
    assert not np.array_equal(unreduced, util.gt_img(o.miller_loop(o.g2_precompute(Qa), P)))
    assert np.array_equal(emu.gt_op(6, unreduced), gt[0])


@pytest.mark.parametrize("variant", ["product", "refchain"])
def test_duo_line_schedule_equals_solo(pairs, variant):
    """duo.cuh: the lane-pair line kernel (one Fq2 component per lane) stores the same lines as the one-thread version,
    for the NAF schedule and for the reference's binary chain (102 lines, src/groups/mod.rs:557-588)."""
    g1, g2, _ = pairs
    e1, e2 = util.edge_case_pairs()
    for a, b in [(g1[0], g2[0]), (g1[1], g2[1]), (g1[2], g2[3]), (e1[0], e2[0]), (e1[6], e2[6]), (e1[1], e2[1])]:
        finite, L, _, _ = emu.lines(a, b, variant)
        f2, Ld = emu.lines_duo(a, b, variant)
        assert finite == f2
        if finite:
            assert np.array_equal(L, Ld)


def test_hexad_fq12_ops(pairs):
    _, _, gt = pairs
    a, b = gt[0], gt[1]
    assert np.array_equal(emu.gt_op(0, a, b), cref.fq12_mul(a[None], b[None])[0])
    assert np.array_equal(emu.gt_op(1, a), cref.fq12_sqr(a[None])[0])
    assert np.array_equal(emu.gt_op(2, a), util.gt_img(o.fq12_cyclotomic_squared(util.img_gt(a))))
    assert np.array_equal(emu.gt_op(3, a), cref.fq12_inv(a[None])[0])
    assert np.array_equal(emu.gt_op(7, a), util.gt_img(o.fq12_conj(util.img_gt(a))))
    for p in (1, 2, 3):
        assert np.array_equal(emu.gt_op(4, a, arg=p), cref.fq12_frobenius(a[None], p)[0])
    assert np.array_equal(emu.gt_op(5, a), cref.fq12_exp_by_neg_z(a[None])[0])
    # generic (non-cyclotomic) element: the reference's fq12_test_vector start value
    f = util.load_json("fq12_kat.json")
    s = util.gt_img(o.fq12_from_flat(f["vector_start"]))
    nxt = s
    for _ in range(5):
        nxt = emu.gt_op(0, nxt, s)
    nxt = emu.gt_op(1, nxt)
    want = util.img_gt(s)
    acc = want
    for _ in range(5):
        acc = o.fq12_mul(acc, want)
    assert np.array_equal(nxt, util.gt_img(o.fq12_sqr(acc)))
    assert np.array_equal(emu.gt_op(3, s), util.gt_img(o.fq12_inv(want)))
    k = o.synth_scalar(9, 0)
    assert np.array_equal(emu.gt_op(8, a, _w(k)), cref.gt_pow_batch(a[None], util.fr_img(k)[None])[0])
    for kk in (k, 0, 1, 2, 3, o.R_ORDER - 1):  # cyclotomic fixed-window pow == generic pow on pairing values
        assert np.array_equal(emu.gt_op(9, a, _w(kk)), cref.gt_pow_batch(a[None], util.fr_img(kk)[None])[0]), kk


def test_pairing_kat_and_random(pairs):
    g1, g2, gt = pairs
    kat = util.load_json("pairing_kat.json")
    kg1 = cref.g1_mul_batch(cref.g1_generator(), util.fr_img(int(kat["k1"]))[None])[0]
    kg2 = cref.g2_mul_batch(cref.g2_generator(), util.fr_img(int(kat["k2"]))[None])[0]
    assert np.array_equal(emu.pairing(kg1, kg2), util.gt_img(o.fq12_from_flat(kat["reduced_pairing"])))
    for i in range(len(g1)):
        assert np.array_equal(emu.pairing(g1[i], g2[i]), gt[i])


def test_pairing_edge_cases():
    e1, e2 = util.edge_case_pairs()
    egt = cref.pairing_batch(e1, e2)
    for i in range(len(e1)):
        assert np.array_equal(emu.pairing(e1[i], e2[i]), egt[i]), i


def test_hexad_ops_extreme_coefficients():
    """Bound checks of the lazy arithmetic (512-bit sums that wrap, lazy 9-limb post-processing): all-(q-1), all-zero,
    identity and mixed coefficient patterns through the dense product, the squaring and the Granger-Scott squaring
    (a polynomial map: the reference evaluates it on non-cyclotomic inputs too, src/fields/mod.rs:171-201)."""
    import random
    rng = random.Random(5)
    qm1 = o.Q - 1
    pats = [[qm1] * 12, [0] * 12, [1] + [0] * 11, [qm1, 0] * 6, [0, qm1] * 6, [qm1] * 6 + [0] * 6,
            [rng.choice((0, 1, qm1, qm1 - 1, o.Q // 2)) for _ in range(12)], [rng.randrange(o.Q) for _ in range(12)]]
    elems = [o.fq12_from_flat(p) for p in pats]
    for i, x in enumerate(elems):
        xi = util.gt_img(x)
        assert np.array_equal(emu.gt_op(2, xi), util.gt_img(o.fq12_cyclotomic_squared(x))), ("cyc", i)
        assert np.array_equal(emu.gt_op(1, xi), util.gt_img(o.fq12_sqr(x))), ("sqr", i)
        y = elems[(i + 3) % len(elems)]
        assert np.array_equal(emu.gt_op(0, xi, util.gt_img(y)), util.gt_img(o.fq12_mul(x, y))), ("mul", i)


def test_f52_product_path_matches_integer_path():
    """f52.cuh (5 x 52-bit limbs on the FP64 pipe, Montgomery radix 2^256 by 4 x 52 + 48 bit rounds) against the oracle's
    Montgomery product, extremes included."""
    import random
    rng = random.Random(52)
    vals = [0, 1, 2, o.Q - 1, o.Q - 2, (1 << 254) % o.Q, (1 << 52) - 1, 1 << 52, (1 << 208) + 1, o.Q // 2]
    vals += [rng.randrange(o.Q) for _ in range(60)]
    rinv = pow(1 << 256, -1, o.Q)
    w = lambda v: util.words(v.to_bytes(32, "little"))
    for i, a in enumerate(vals):
        for b in (vals[(3 * i + 1) % len(vals)], vals[(7 * i + 5) % len(vals)], a):
            got = emu.f52_mul(w(a), w(b))
            assert int.from_bytes(got.tobytes(), "little") == a * b * rinv % o.Q, (a, b)


def test_dedicated_squaring_matches_multiplication():
    """fp_sqr (36-IMAD triangle + diagonal, fp.cuh wide_sqr) == fp_mul(a, a) for Fq and Fr, extremes included."""
    import random
    rng = random.Random(108)
    for which, mod in ((0, o.Q), (1, o.R_ORDER)):
        vals = [0, 1, 2, mod - 1, mod - 2, (1 << 253) % mod, (1 << 32) - 1, (1 << 224) + (1 << 32) - 1, mod // 2] + [rng.randrange(mod) for _ in range(100)]
        for v in vals:
            w = util.words(v.to_bytes(32, "little"))
            assert np.array_equal(emu.fp_op(8, which, w), emu.fp_op(0, which, w, w)), (which, v)
