"""Batch entry points on numpy byte images + thin value classes mirroring reference src/lib.rs.

Byte images are the crate's #[repr(C)] layouts as numpy uint64 arrays:
  Fr [n,4]   G1 [n,12]   G2 [n,24]   Gt [n,48]      (Montgomery form, LE limbs)
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

FR_WORDS, G1_WORDS, G2_WORDS, GT_WORDS = 4, 12, 24, 48


def _arr(a, words):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.shape[-1] != words:
        raise ValueError("expected [n,%d] uint64 image, got %s" % (words, a.shape))
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _binary(fn_name, a, wa, b, wb, wo, device):
    lib = _lib.ensure(device)
    a, b = _arr(a, wa), _arr(b, wb)
    if len(a) != len(b):
        raise ValueError("batch length mismatch")
    out = np.empty((len(a), wo), dtype=np.uint64)
    _lib.check(getattr(lib, fn_name)(_p(a), _p(b), _p(out), ctypes.c_size_t(len(a))))
    return out


def pairing_batch(g1, g2, device=None) -> np.ndarray:
    """[n] pairings e(g1[i], g2[i]) -> Gt images.  reference bn::pairing, src/lib.rs:181-183."""
    return _binary("bn_b200_pairing_batch", g1, G1_WORDS, g2, G2_WORDS, GT_WORDS, device)


def pairing_pow_batch(g1, g2, fr, device=None) -> np.ndarray:
    """pairing(g1[i], g2[i]).pow(fr[i]) fused in one pass (reference examples/joux.rs:19-21)."""
    lib = _lib.ensure(device)
    g1, g2, fr = _arr(g1, G1_WORDS), _arr(g2, G2_WORDS), _arr(fr, FR_WORDS)
    if not (len(g1) == len(g2) == len(fr)):
        raise ValueError("batch length mismatch")
    out = np.empty((len(g1), GT_WORDS), dtype=np.uint64)
    _lib.check(lib.bn_b200_pairing_pow_batch(_p(g1), _p(g2), _p(fr), _p(out), ctypes.c_size_t(len(g1))))
    return out


def g1_mul_batch(g1, fr, device=None) -> np.ndarray:
    """g1[i] * fr[i] (Jacobian, un-normalised like the crate).  reference src/lib.rs:116-120."""
    return _binary("bn_b200_g1_mul_batch", g1, G1_WORDS, fr, FR_WORDS, G1_WORDS, device)


def g2_mul_batch(g2, fr, device=None) -> np.ndarray:
    """reference src/lib.rs:159-163."""
    return _binary("bn_b200_g2_mul_batch", g2, G2_WORDS, fr, FR_WORDS, G2_WORDS, device)


def gt_pow_batch(gt, fr, device=None) -> np.ndarray:
    """gt[i].pow(fr[i]).  reference Gt::pow, src/lib.rs:171."""
    return _binary("bn_b200_gt_pow_batch", gt, GT_WORDS, fr, FR_WORDS, GT_WORDS, device)


def gt_mul_batch(a, b, device=None) -> np.ndarray:
    """a[i] * b[i].  reference src/lib.rs:175-179."""
    return _binary("bn_b200_gt_mul_batch", a, GT_WORDS, b, GT_WORDS, GT_WORDS, device)


def gt_inv_batch(a, device=None) -> np.ndarray:
    """a[i].inverse().  reference Gt::inverse, src/lib.rs:172."""
    lib = _lib.ensure(device)
    a = _arr(a, GT_WORDS)
    out = np.empty_like(a)
    _lib.check(lib.bn_b200_gt_inv_batch(_p(a), _p(out), ctypes.c_size_t(len(a))))
    return out


FR_OPS = {"mul": 0, "add": 1, "sub": 2, "neg": 3, "inverse": 4}


def fr_op_batch(op: str, a, b=None, device=None) -> np.ndarray:
    """Batched Fr arithmetic (reference src/lib.rs:19-54): op in mul/add/sub/neg/inverse."""
    lib = _lib.ensure(device)
    a = _arr(a, FR_WORDS)
    b = _arr(b, FR_WORDS) if b is not None else None
    out = np.empty_like(a)
    _lib.check(lib.bn_b200_fr_op_batch(FR_OPS[op], _p(a), _p(b) if b is not None else None, _p(out),
                                       ctypes.c_size_t(len(a))))
    return out


def g1_normalize_batch(g1, device=None) -> np.ndarray:
    """Group::normalize for G1 (reference src/lib.rs:88-95)."""
    lib = _lib.ensure(device)
    g1 = _arr(g1, G1_WORDS)
    out = np.empty_like(g1)
    _lib.check(lib.bn_b200_g1_normalize_batch(_p(g1), _p(out), ctypes.c_size_t(len(g1))))
    return out


def g2_normalize_batch(g2, device=None) -> np.ndarray:
    """Group::normalize for G2 (reference src/lib.rs:131-138)."""
    lib = _lib.ensure(device)
    g2 = _arr(g2, G2_WORDS)
    out = np.empty_like(g2)
    _lib.check(lib.bn_b200_g2_normalize_batch(_p(g2), _p(out), ctypes.c_size_t(len(g2))))
    return out


def g1_check_batch(g1, device=None) -> np.ndarray:
    """On-curve check of decoded G1 points (reference AffineG::decode, src/groups/mod.rs:178-205) -> bool[n]."""
    lib = _lib.ensure(device)
    g1 = _arr(g1, G1_WORDS)
    ok = np.zeros(len(g1), dtype=np.uint8)
    _lib.check(lib.bn_b200_g1_check_batch(_p(g1), ok.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(g1))))
    return ok.astype(bool)


def g2_check_batch(g2, device=None) -> np.ndarray:
    """On-curve + order-r subgroup check of decoded G2 points (src/groups/mod.rs:183-195) -> bool[n]."""
    lib = _lib.ensure(device)
    g2 = _arr(g2, G2_WORDS)
    ok = np.zeros(len(g2), dtype=np.uint8)
    _lib.check(lib.bn_b200_g2_check_batch(_p(g2), ok.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(g2))))
    return ok.astype(bool)


WIRE_BYTES = {"g1": 65, "g2": 129, "fr": 32}
_WIRE_WORDS = {"g1": G1_WORDS, "g2": G2_WORDS, "fr": FR_WORDS}
WIRE_ERRORS = {1: "invalid leading byte for uncompressed group element", 2: "integer is not less than modulus",
               3: "point is not on the curve", 4: "point is not in the subgroup"}


def encode_batch(kind: str, img, device=None) -> np.ndarray:
    """Wire records [n, 65 | 129 | 32] uint8 of G1 / G2 / Fr images (reference RustcEncodable impls:
    src/groups/mod.rs:143-163, src/fields/fq2.rs:31-40, src/fields/fp.rs:24-29).  Infinity = 0x00 + zero padding."""
    lib = _lib.ensure(device)
    img = _arr(img, _WIRE_WORDS[kind])
    out = np.zeros((len(img), WIRE_BYTES[kind]), dtype=np.uint8)
    _lib.check(getattr(lib, "bn_b200_%s_encode_batch" % kind)(_p(img), _p(out), ctypes.c_size_t(len(img))))
    return out


def decode_batch(kind: str, records, device=None):
    """Wire records -> (images, status[n]); status 0 = ok, else a key of WIRE_ERRORS (reference RustcDecodable impls:
    src/groups/mod.rs:178-205, src/fields/fq2.rs:42-53, src/fields/fp.rs:31-36)."""
    lib = _lib.ensure(device)
    rec = np.ascontiguousarray(records, dtype=np.uint8)
    if rec.ndim != 2 or rec.shape[1] != WIRE_BYTES[kind]:
        raise ValueError("expected [n,%d] uint8 records, got %s" % (WIRE_BYTES[kind], rec.shape))
    out = np.empty((len(rec), _WIRE_WORDS[kind]), dtype=np.uint64)
    status = np.zeros(len(rec), dtype=np.uint8)
    _lib.check(getattr(lib, "bn_b200_%s_decode_batch" % kind)(_p(rec), _p(out), _p(status), ctypes.c_size_t(len(rec))))
    return out, status


def to_wire(kind: str, record: np.ndarray) -> bytes:
    """The exact reference byte string of one record (infinity is the single byte 0x00)."""
    b = bytes(record)
    return b[:1] if kind != "fr" and b[0] == 0 else b


def from_wire(kind: str, data: bytes) -> np.ndarray:
    """One reference byte string -> fixed-stride record (zero padded)."""
    r = np.zeros(WIRE_BYTES[kind], dtype=np.uint8)
    d = np.frombuffer(data, dtype=np.uint8)
    if len(d) > len(r):
        raise ValueError("record too long")
    r[:len(d)] = d
    return r


def fq_mul_chain(a, b, iters: int, device=None) -> np.ndarray:
    """x <- x*b (Montgomery mod q) `iters` times per element (BASELINE config 2)."""
    lib = _lib.ensure(device)
    a, b = _arr(a, 4), _arr(b, 4)
    out = np.empty_like(a)
    _lib.check(lib.bn_b200_fq_mul_chain(_p(a), _p(b), _p(out), ctypes.c_size_t(len(a)), ctypes.c_uint32(iters)))
    return out


GROUP_OPS = {"add": 0, "sub": 1, "neg": 2, "double": 3}


def _group_op(kind, words, op, a, b, device):
    lib = _lib.ensure(device)
    a = _arr(a, words)
    b = _arr(b, words) if b is not None else None
    if b is not None and len(a) != len(b):
        raise ValueError("batch length mismatch")
    out = np.empty_like(a)
    _lib.check(getattr(lib, "bn_b200_%s_op_batch" % kind)(GROUP_OPS[op], _p(a), _p(b) if b is not None else None, _p(out),
                                                         ctypes.c_size_t(len(a))))
    return out


def g1_op_batch(op: str, a, b=None, device=None) -> np.ndarray:
    """Batched group law on G1: op in add / sub / neg / double (reference impl Add / Sub / Neg for G1, src/lib.rs:97-114;
    G::double, src/groups/mod.rs:228-247).  Un-normalised Jacobian output, limb-identical to the crate's."""
    return _group_op("g1", G1_WORDS, op, a, b, device)


def g2_op_batch(op: str, a, b=None, device=None) -> np.ndarray:
    """Batched group law on G2 (reference src/lib.rs:140-157)."""
    return _group_op("g2", G2_WORDS, op, a, b, device)


def _group_eq(kind, words, a, b, device):
    lib = _lib.ensure(device)
    a, b = _arr(a, words), _arr(b, words)
    if len(a) != len(b):
        raise ValueError("batch length mismatch")
    eq = np.zeros(len(a), dtype=np.uint8)
    _lib.check(getattr(lib, "bn_b200_%s_eq_batch" % kind)(_p(a), _p(b), _p(eq), ctypes.c_size_t(len(a))))
    return eq.astype(bool)


def g1_eq_batch(a, b, device=None) -> np.ndarray:
    """PartialEq for G1: projective equality (reference src/groups/mod.rs:83-109) -> bool[n]."""
    return _group_eq("g1", G1_WORDS, a, b, device)


def g2_eq_batch(a, b, device=None) -> np.ndarray:
    return _group_eq("g2", G2_WORDS, a, b, device)


def fr_pow_batch(a, e, device=None) -> np.ndarray:
    """a[i].pow(e[i]) in Fr (reference Fr::pow, src/lib.rs:24)."""
    return _binary("bn_b200_fr_pow_batch", a, FR_WORDS, e, FR_WORDS, FR_WORDS, device)


def gt_exp_by_neg_z_batch(a, device=None) -> np.ndarray:
    """Fq12::exp_by_neg_z with the reference's literal chain (src/fields/fq12.rs:97-101), valid for any Fq12 input."""
    lib = _lib.ensure(device)
    a = _arr(a, GT_WORDS)
    out = np.empty_like(a)
    _lib.check(lib.bn_b200_gt_exp_by_neg_z_batch(_p(a), _p(out), ctypes.c_size_t(len(a))))
    return out


def fq_sqr_chain(a, iters: int, device=None) -> np.ndarray:
    """x <- x^2 (Montgomery mod q) `iters` times per element (dedicated squaring)."""
    lib = _lib.ensure(device)
    a = _arr(a, 4)
    out = np.empty_like(a)
    _lib.check(lib.bn_b200_fq_sqr_chain(_p(a), _p(out), ctypes.c_size_t(len(a)), ctypes.c_uint32(iters)))
    return out


# Montgomery images of the constants the value classes need (reference src/fields/fp.rs:161-177, src/groups/mod.rs:356-390)
_Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def _mont(x, mod):
    return np.frombuffer((x * (1 << 256) % mod).to_bytes(32, "little"), dtype="<u8").astype(np.uint64)


_G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
            11559732032986387107991004021392285783925812861821192530917403151452391805634),
           (8495653923123431417604973247489272438418190587263600148770280649306958101930,
            4082367875863433681332203403145435568316851327593401208105741076214120093531))


class _Img:
    WORDS = 0

    def __init__(self, img):
        self.img = np.ascontiguousarray(img, dtype=np.uint64).reshape(self.WORDS)

    def __eq__(self, other):  # limb equality (Gt / Fr are canonical; for G1/G2 this is stricter than the crate's projective eq)
        return type(self) is type(other) and bool(np.array_equal(self.img, other.img))

    def __hash__(self):
        return hash(self.img.tobytes())


class Fr(_Img):
    """reference src/lib.rs:15-54 (the scalar is carried as its Montgomery image)."""
    WORDS = FR_WORDS

    @classmethod
    def zero(cls):
        return cls(np.zeros(4, dtype=np.uint64))

    @classmethod
    def one(cls):
        return cls(_mont(1, _R))

    @classmethod
    def from_int(cls, v: int):
        """Mirror of Fr::from_str on a decimal value (src/fields/fp.rs:39-59: wraps mod r)."""
        return cls(_mont(v % _R, _R))

    def is_zero(self) -> bool:
        return not self.img.any()

    def _op(self, op, other=None):
        return Fr(fr_op_batch(op, self.img[None], other.img[None] if other is not None else None)[0])

    def __add__(self, o):
        return self._op("add", o)

    def __sub__(self, o):
        return self._op("sub", o)

    def __mul__(self, o):
        return self._op("mul", o)

    def __neg__(self):
        return self._op("neg")

    def inverse(self):
        """None for zero, like the crate (src/lib.rs:26)."""
        return None if self.is_zero() else self._op("inverse")

    def pow(self, e: "Fr") -> "Fr":
        return Fr(fr_pow_batch(self.img[None], e.img[None])[0])


class _GroupMixin:
    """reference trait Group, src/lib.rs:56-77: zero / one / is_zero / normalize + Add / Sub / Neg / Mul<Fr>."""
    _kind = ""

    def _gop(self, op, other=None):
        return type(self)(_group_op(self._kind, self.WORDS, op, self.img[None], other.img[None] if other is not None else None, None)[0])

    def __add__(self, o):
        return self._gop("add", o)

    def __sub__(self, o):
        return self._gop("sub", o)

    def __neg__(self):
        return self._gop("neg")

    def double(self):
        return self._gop("double")

    def is_zero(self) -> bool:
        return not self.img[2 * self.WORDS // 3:].any()

    def __eq__(self, other):  # projective equality, like the crate's derived PartialEq over groups::G
        return type(self) is type(other) and bool(_group_eq(self._kind, self.WORDS, self.img[None], other.img[None], None)[0])

    __hash__ = None


class G1(_GroupMixin, _Img):
    """reference src/lib.rs:79-120."""
    WORDS = G1_WORDS
    _kind = "g1"

    @classmethod
    def zero(cls):
        return cls(np.concatenate([np.zeros(4, np.uint64), _mont(1, _Q), np.zeros(4, np.uint64)]))

    @classmethod
    def one(cls):
        return cls(np.concatenate([_mont(1, _Q), _mont(2, _Q), _mont(1, _Q)]))

    def normalize(self):
        self.img = g1_normalize_batch(self.img[None])[0]

    def __mul__(self, k: Fr) -> "G1":
        return G1(g1_mul_batch(self.img[None], k.img[None])[0])


class G2(_GroupMixin, _Img):
    """reference src/lib.rs:122-163."""
    WORDS = G2_WORDS
    _kind = "g2"

    @classmethod
    def zero(cls):
        z = np.zeros(8, np.uint64)
        return cls(np.concatenate([z, _mont(1, _Q), np.zeros(4, np.uint64), z]))

    @classmethod
    def one(cls):
        (x0, x1), (y0, y1) = _G2_GEN
        return cls(np.concatenate([_mont(x0, _Q), _mont(x1, _Q), _mont(y0, _Q), _mont(y1, _Q), _mont(1, _Q), np.zeros(4, np.uint64)]))

    def normalize(self):
        self.img = g2_normalize_batch(self.img[None])[0]

    def __mul__(self, k: Fr) -> "G2":
        return G2(g2_mul_batch(self.img[None], k.img[None])[0])


class Gt(_Img):
    """reference src/lib.rs:165-179."""
    WORDS = GT_WORDS

    @classmethod
    def one(cls):
        return cls(np.concatenate([_mont(1, _Q), np.zeros(44, np.uint64)]))

    def pow(self, k: Fr) -> "Gt":
        return Gt(gt_pow_batch(self.img[None], k.img[None])[0])

    def __mul__(self, other: "Gt") -> "Gt":
        return Gt(gt_mul_batch(self.img[None], other.img[None])[0])

    def inverse(self) -> "Gt":
        return Gt(gt_inv_batch(self.img[None])[0])


def pairing(p: G1, q: G2) -> Gt:
    """reference src/lib.rs:181-183."""
    return Gt(pairing_batch(p.img[None], q.img[None])[0])
