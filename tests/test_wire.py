"""Row f-3: the crate's wire format (bn_b200/csrc/wire.cuh).  CPU tests run the same header through the host emulator;
GPU tests go through the C ABI.  Pinned on the reference's own serialization vectors (tests/golden/, extracted from
reference tests/serialization.rs) and on the big-int oracle's encode / decode (oracle/bn_oracle.py:765-827)."""
import numpy as np
import pytest

from oracle import bn_oracle as o
from oracle import cref
from tests import emu, util


def _rec(kind, hx):
    n = {"g1": 65, "g2": 129, "fr": 32}[kind]
    r = np.zeros(n, dtype=np.uint8)
    b = np.frombuffer(bytes.fromhex(hx), dtype=np.uint8)
    r[:len(b)] = b
    return r


def _oracle_status(kind, rec):
    """Status the reference's decode would give (0 ok, 1 tag, 2 range, 3 curve, 4 subgroup), from the big-int oracle."""
    b = bytes(rec)
    if kind == "fr":
        return 0 if int.from_bytes(b, "big") < o.R_ORDER else 2
    if b[0] == 0:
        return 0
    if b[0] != 4:
        return 1
    try:
        (o.decode_g1 if kind == "g1" else o.decode_g2)(b)
        return 0
    except ValueError as e:
        msg = str(e)
        return 2 if "modulus" in msg else 3 if "curve" in msg else 4


def _bad_records(kind, good, rng):
    """Mutations of valid records: wrong tag, coordinate >= modulus, off-curve, and (G2) on-curve but outside the subgroup."""
    out = []
    r = good.copy(); r[0] = 0x23; out.append(r)
    r = good.copy(); r[1:33] = 0xFF; out.append(r)                       # x >= q (G1) / x >= q^2 (G2)
    r = good.copy(); r[-1] ^= 1; out.append(r)                           # off the curve (or, rarely, not reduced)
    r = good.copy(); r[0] = 0; r[1:] = rng.integers(0, 256, len(r) - 1); out.append(r)   # infinity: only byte 0 is read
    if kind == "g1":
        r = good.copy(); r[33:65] = np.frombuffer(o.Q.to_bytes(32, "big"), dtype=np.uint8); out.append(r)  # y == q exactly
    else:
        # a point of the twist outside the order-r subgroup: pick x, solve y^2 = x^3 + b' (cofactor is huge, so almost surely outside)
        for x0 in range(1, 50):
            x = (x0, 1)
            rhs = o.fq2_add(o.fq2_mul(o.fq2_sqr(x), x), o.G2_B)
            y = _fq2_sqrt(rhs)
            if y is not None:
                out.append(np.concatenate([[4], np.frombuffer((x[1] * o.Q + x[0]).to_bytes(64, "big") + (y[1] * o.Q + y[0]).to_bytes(64, "big"), dtype=np.uint8)]).astype(np.uint8))
                break
    return out


def _fq2_sqrt(a):
    """square root in Fq2 = Fq[i]/(i^2+1) (q = 3 mod 4), or None"""
    q = o.Q
    if a == (0, 0):
        return (0, 0)
    a1 = o.fq2_pow(a, (q - 3) // 4)
    alpha = o.fq2_mul(o.fq2_mul(a1, a1), a)
    x0 = o.fq2_mul(a1, a)
    if alpha == ((q - 1) % q, 0):
        r = (-x0[1] % q, x0[0])
    else:
        b = o.fq2_pow(o.fq2_add(alpha, (1, 0)), (q - 1) // 2)
        r = o.fq2_mul(b, x0)
    return r if o.fq2_sqr(r) == a else None


def _cases(kind, limit):
    lines = util.load_vectors(kind + "_vectors.txt", limit)
    return [_rec(kind, h) for h in lines]


# ------------------------------------------------------------------ CPU: wire.cuh through the host emulator
@pytest.mark.parametrize("kind,limit", [("g1", 60), ("g2", 12), ("fr", 200)])
def test_emu_wire_roundtrip_golden(kind, limit):
    for rec in _cases(kind, limit):
        img, st = emu.wire_decode(kind, rec)
        assert st == 0
        if kind == "g1":
            assert o.g_eq(o.FQ, util.img_g1(img), o.decode_g1(bytes(rec)))
        elif kind == "g2":
            assert o.g_eq(o.FQ2, util.img_g2(img), o.decode_g2(bytes(rec)))
        else:
            assert np.array_equal(img, util.fr_img(int.from_bytes(bytes(rec), "big")))
        assert np.array_equal(emu.wire_encode(kind, img), rec)


def test_emu_wire_encode_normalises_and_infinity():
    g1, g2 = util.synth_pairs(0xB2000007, 3)
    for kind, imgs, F, dec, enc in (("g1", g1, o.FQ, util.img_g1, o.encode_g1), ("g2", g2, o.FQ2, util.img_g2, o.encode_g2)):
        for im in imgs:  # Jacobian, z != 1
            want = enc(dec(im))
            assert bytes(emu.wire_encode(kind, im)) == want
    e1, e2 = util.edge_case_pairs()
    assert emu.wire_encode("g1", e1[1])[0] == 0 and not emu.wire_encode("g1", e1[1]).any()
    assert emu.wire_encode("g2", e2[2])[0] == 0 and not emu.wire_encode("g2", e2[5]).any()


@pytest.mark.parametrize("kind", ["g1", "g2"])
def test_emu_wire_decode_rejections(kind):
    rng = np.random.default_rng(7)
    good = _cases(kind, 2)[1]
    bad = _bad_records(kind, good, rng)
    seen = set()
    for rec in bad:
        img, st = emu.wire_decode(kind, rec)
        assert st == _oracle_status(kind, rec), bytes(rec).hex()
        seen.add(st)
        if st != 0 or rec[0] == 0:
            zero = util.g1_img(o.g_zero(o.FQ)) if kind == "g1" else util.g2_img(o.g_zero(o.FQ2))
            assert np.array_equal(img, zero)
    assert {0, 1, 2, 3} <= seen and (kind == "g1" or 4 in seen)
    for k, hx in util.load_json("wire_edge_cases.json")["cases"]:
        if k.lower() == kind:
            rec = _rec(kind, hx)
            _, st = emu.wire_decode(kind, rec)
            assert (st == 0) == (hx == "00") and st == _oracle_status(kind, rec)


def test_emu_fr_decode_range():
    for v, want in ((o.R_ORDER - 1, 0), (o.R_ORDER, 2), ((1 << 256) - 1, 2), (0, 0)):
        _, st = emu.wire_decode("fr", np.frombuffer(v.to_bytes(32, "big"), dtype=np.uint8))
        assert st == want


# ------------------------------------------------------------------ GPU: the same through the C ABI
@pytest.fixture(scope="module")
def bn():
    import bn_b200
    bn_b200.init(0)
    return bn_b200


@pytest.mark.gpu
@pytest.mark.parametrize("kind,limit", [("g1", 2000), ("g2", 400), ("fr", 10000)])
def test_gpu_wire_golden_vectors(bn, kind, limit):
    """decode every stored reference vector, compare with the oracle's value, encode back: identical bytes."""
    from bn_b200 import api
    recs = np.stack(_cases(kind, limit))
    imgs, st = api.decode_batch(kind, recs)
    assert not st.any()
    for i in range(0, len(recs), max(1, len(recs) // 50)):  # spot-check values against the big-int oracle
        if kind == "g1":
            assert o.g_eq(o.FQ, util.img_g1(imgs[i]), o.decode_g1(bytes(recs[i])))
        elif kind == "g2":
            p = util.img_g2(imgs[i])
            x, y = int.from_bytes(bytes(recs[i][1:65]), "big"), int.from_bytes(bytes(recs[i][65:]), "big")
            assert p == ((x % o.Q, x // o.Q), (y % o.Q, y // o.Q), o.FQ2_ONE)
        else:
            assert np.array_equal(imgs[i], util.fr_img(int.from_bytes(bytes(recs[i]), "big")))
    assert np.array_equal(api.encode_batch(kind, imgs), recs)


@pytest.mark.gpu
def test_gpu_wire_encode_jacobian_inputs(bn):
    from bn_b200 import api
    g1, g2 = util.synth_pairs(0xB2000008, 64)
    e1, e2 = util.edge_case_pairs()
    g1, g2 = np.concatenate([g1, e1]), np.concatenate([g2, e2])
    r1, r2 = api.encode_batch("g1", g1), api.encode_batch("g2", g2)
    n1, n2 = cref.g1_normalize(g1), cref.g2_normalize(g2)
    for i in range(len(g1)):
        assert api.to_wire("g1", r1[i]) == o.encode_g1(util.img_g1(n1[i])), i
        assert api.to_wire("g2", r2[i]) == o.encode_g2(util.img_g2(n2[i])), i
    # round trip: decode(encode(p)) is the normalised point
    d1, s1 = api.decode_batch("g1", r1)
    d2, s2 = api.decode_batch("g2", r2)
    assert not s1.any() and not s2.any()
    for i in range(len(g1)):
        assert o.g_eq(o.FQ, util.img_g1(d1[i]), util.img_g1(g1[i])) and o.g_eq(o.FQ2, util.img_g2(d2[i]), util.img_g2(g2[i]))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["g1", "g2"])
def test_gpu_wire_decode_rejections(bn, kind):
    from bn_b200 import api
    rng = np.random.default_rng(11)
    good = _cases(kind, 8)
    recs = []
    for g in good:
        recs += _bad_records(kind, g, rng)
    recs += [_rec(kind, hx) for k, hx in util.load_json("wire_edge_cases.json")["cases"] if k.lower() == kind]
    recs = np.stack(recs)
    imgs, st = api.decode_batch(kind, recs)
    want = np.array([_oracle_status(kind, r) for r in recs], dtype=np.uint8)
    assert np.array_equal(st, want)
    assert {0, 1, 2, 3} <= set(st.tolist()) and (kind == "g1" or 4 in set(st.tolist()))
    zero = util.g1_img(o.g_zero(o.FQ)) if kind == "g1" else util.g2_img(o.g_zero(o.FQ2))
    for i in range(len(recs)):
        if st[i] != 0 or recs[i][0] == 0:
            assert np.array_equal(imgs[i], zero)


@pytest.mark.gpu
def test_gpu_fr_wire(bn):
    from bn_b200 import api
    vals = [0, 1, o.R_ORDER - 1, o.R_ORDER, o.R_ORDER + 5, (1 << 256) - 1] + [o.synth_scalar(0xB2000009, i) for i in range(100)]
    recs = np.stack([np.frombuffer(v.to_bytes(32, "big"), dtype=np.uint8) for v in vals])
    imgs, st = api.decode_batch("fr", recs)
    assert np.array_equal(st, np.array([0 if v < o.R_ORDER else 2 for v in vals], dtype=np.uint8))
    ok = st == 0
    assert np.array_equal(imgs[ok], np.stack([util.fr_img(v) for v in vals if v < o.R_ORDER]))
    assert np.array_equal(api.encode_batch("fr", imgs[ok]), recs[ok])


def test_emu_wire_decode_fuzz():
    """Random mutations of valid records and fully random records: status and decoded value must agree with the big-int
    oracle's decode (the reference's checks in the reference's order)."""
    rng = np.random.default_rng(20261017)
    for kind, limit, rounds in (("g1", 6, 60), ("g2", 3, 14), ("fr", 4, 60)):
        good = _cases(kind, limit)
        for it in range(rounds):
            rec = good[it % len(good)].copy()
            mode = it % 4
            if mode == 0:
                rec[int(rng.integers(0, len(rec)))] ^= np.uint8(1 << int(rng.integers(0, 8)))   # one flipped bit
            elif mode == 1:
                rec[1 if kind != "fr" else 0] = np.uint8(rng.integers(0x30, 0x100))              # a large leading coordinate byte
            elif mode == 2:
                rec[:] = rng.integers(0, 256, len(rec), dtype=np.uint8)                          # noise
                if kind != "fr" and it % 8 == 2:
                    rec[0] = 4
            img, st = emu.wire_decode(kind, rec)
            assert st == _oracle_status(kind, rec), (kind, it, bytes(rec).hex())
            if st == 0 and kind == "fr":
                assert np.array_equal(img, util.fr_img(int.from_bytes(bytes(rec), "big")))
            elif st == 0 and rec[0] == 4:
                dec, F, to = (o.decode_g1, o.FQ, util.img_g1) if kind == "g1" else (o.decode_g2, o.FQ2, util.img_g2)
                assert o.g_eq(F, to(img), dec(bytes(rec)))
