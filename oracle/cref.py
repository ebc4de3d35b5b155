"""ctypes binding for the C oracle oracle/libbn_ref.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  Buffers are numpy uint64 arrays holding the crate's #[repr(C)]
byte images (Montgomery form, 4 LE u64 limbs per field element):
  Fr/Fq [n,4]   G1 [n,12]   G2 [n,24]   Gt [n,48]
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbn_ref.so")
_lib = None

FR_WORDS, G1_WORDS, G2_WORDS, GT_WORDS = 4, 12, 24, 48


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bn_ref.c")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("bn_ref.c", "bn_ref_consts.h") if os.path.exists(os.path.join(_HERE, f)))
    if force or stale:
        if not os.path.exists(os.path.join(_HERE, "bn_ref_consts.h")):
            subprocess.check_call(["python", os.path.join(_HERE, "gen_consts.py")])
        subprocess.check_call(
            ["gcc", "-O3", "-march=native", "-fPIC", "-fvisibility=hidden", "-std=gnu11", "-shared",
             "-o", _SO, src, "-lpthread"], cwd=_HERE)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def _arr(a, words):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, words)
    return a


def pairing_batch(g1, g2, threads: int = 1) -> np.ndarray:
    g1, g2 = _arr(g1, G1_WORDS), _arr(g2, G2_WORDS)
    assert len(g1) == len(g2)
    out = np.zeros((len(g1), GT_WORDS), dtype=np.uint64)
    lib().bn_ref_pairing_batch(_p(g1), _p(g2), _p(out), ctypes.c_size_t(len(g1)), ctypes.c_int(threads))
    return out


def g1_mul_batch(g1, fr, threads: int = 1) -> np.ndarray:
    g1, fr = _arr(g1, G1_WORDS), _arr(fr, FR_WORDS)
    out = np.zeros_like(g1)
    lib().bn_ref_g1_mul_batch(_p(g1), _p(fr), _p(out), ctypes.c_size_t(len(g1)), ctypes.c_int(threads))
    return out


def g2_mul_batch(g2, fr, threads: int = 1) -> np.ndarray:
    g2, fr = _arr(g2, G2_WORDS), _arr(fr, FR_WORDS)
    out = np.zeros_like(g2)
    lib().bn_ref_g2_mul_batch(_p(g2), _p(fr), _p(out), ctypes.c_size_t(len(g2)), ctypes.c_int(threads))
    return out


def gt_pow_batch(gt, fr, threads: int = 1) -> np.ndarray:
    gt, fr = _arr(gt, GT_WORDS), _arr(fr, FR_WORDS)
    out = np.zeros_like(gt)
    lib().bn_ref_gt_pow_batch(_p(gt), _p(fr), _p(out), ctypes.c_size_t(len(gt)), ctypes.c_int(threads))
    return out


def gt_mul_batch(a, b, threads: int = 1) -> np.ndarray:
    a, b = _arr(a, GT_WORDS), _arr(b, GT_WORDS)
    out = np.zeros_like(a)
    lib().bn_ref_gt_mul_batch(_p(a), _p(b), _p(out), ctypes.c_size_t(len(a)), ctypes.c_int(threads))
    return out


def fq_mul_chain(a, b, iters: int, threads: int = 1) -> np.ndarray:
    a, b = _arr(a, FR_WORDS), _arr(b, FR_WORDS)
    out = np.zeros_like(a)
    lib().bn_ref_fq_mul_chain(_p(a), _p(b), _p(out), ctypes.c_size_t(len(a)), ctypes.c_uint32(iters),
                              ctypes.c_int(threads))
    return out


def _unary(name, a, words, *extra):
    a = _arr(a, words)
    out = np.zeros_like(a)
    fn = getattr(lib(), name)
    for i in range(len(a)):
        fn(_p(a[i:i + 1]), *extra, _p(out[i:i + 1]))
    return out


def _binary(name, a, b, words):
    a, b = _arr(a, words), _arr(b, words)
    out = np.zeros_like(a)
    fn = getattr(lib(), name)
    for i in range(len(a)):
        fn(_p(a[i:i + 1]), _p(b[i:i + 1]), _p(out[i:i + 1]))
    return out


def g1_add(a, b): return _binary("bn_ref_g1_add", a, b, G1_WORDS)
def g2_add(a, b): return _binary("bn_ref_g2_add", a, b, G2_WORDS)
def g1_double(a): return _unary("bn_ref_g1_double", a, G1_WORDS)
def g2_double(a): return _unary("bn_ref_g2_double", a, G2_WORDS)
def g1_neg(a): return _unary("bn_ref_g1_neg", a, G1_WORDS)
def g2_neg(a): return _unary("bn_ref_g2_neg", a, G2_WORDS)
def g1_normalize(a): return _unary("bn_ref_g1_normalize", a, G1_WORDS)
def g2_normalize(a): return _unary("bn_ref_g2_normalize", a, G2_WORDS)
def fq12_mul(a, b): return _binary("bn_ref_fq12_mul", a, b, GT_WORDS)
def fq12_add(a, b): return _binary("bn_ref_fq12_add", a, b, GT_WORDS)
def fq12_sub(a, b): return _binary("bn_ref_fq12_sub", a, b, GT_WORDS)
def fq12_sqr(a): return _unary("bn_ref_fq12_sqr", a, GT_WORDS)
def fq12_neg(a): return _unary("bn_ref_fq12_neg", a, GT_WORDS)
def fq12_inv(a): return _unary("bn_ref_fq12_inv", a, GT_WORDS)
def fq12_exp_by_neg_z(a): return _unary("bn_ref_fq12_exp_by_neg_z", a, GT_WORDS)
def final_exponentiation(a): return _unary("bn_ref_final_exponentiation", a, GT_WORDS)


def fq12_frobenius(a, power: int):
    a = _arr(a, GT_WORDS)
    out = np.zeros_like(a)
    for i in range(len(a)):
        lib().bn_ref_fq12_frobenius(_p(a[i:i + 1]), ctypes.c_int(power), _p(out[i:i + 1]))
    return out


def g1_generator() -> np.ndarray:
    out = np.zeros((1, G1_WORDS), dtype=np.uint64)
    lib().bn_ref_g1_generator(_p(out))
    return out


def g2_generator() -> np.ndarray:
    out = np.zeros((1, G2_WORDS), dtype=np.uint64)
    lib().bn_ref_g2_generator(_p(out))
    return out


def g2_precompute(q_affine_xy) -> np.ndarray:
    """q_affine_xy: [8] words (x.c0,x.c1,y.c0,y.c1 limbs) -> [102, 24] (ell_0, ell_vw, ell_vv)."""
    q = _arr(q_affine_xy, 16)
    out = np.zeros((102, 24), dtype=np.uint64)
    n = lib().bn_ref_g2_precompute(_p(q[:, :8].copy()), _p(q[:, 8:].copy()), _p(out))
    assert n == 102
    return out


def miller_loop(coeffs, p_affine_xy) -> np.ndarray:
    c = _arr(coeffs, 24)
    p = _arr(p_affine_xy, 8)
    out = np.zeros((1, GT_WORDS), dtype=np.uint64)
    lib().bn_ref_miller_loop(_p(c), _p(p[:, :4].copy()), _p(p[:, 4:].copy()), _p(out))
    return out


def fp_op(op: str, which: int, a, b=None) -> np.ndarray:
    a = _arr(a, 4)
    out = np.zeros_like(a)
    fn = getattr(lib(), "bn_ref_fp_" + op)
    if b is not None:
        b = _arr(b, 4)
    for i in range(len(a)):
        if b is None:
            fn(ctypes.c_int(which), _p(a[i:i + 1]), _p(out[i:i + 1]))
        else:
            fn(ctypes.c_int(which), _p(a[i:i + 1]), _p(b[i:i + 1]), _p(out[i:i + 1]))
    return out
