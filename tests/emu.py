"""ctypes driver for tests/host_emu/libemu.so (host emulation of the kernel arithmetic; test infra only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu")
_libs = {}


def lib(variant: str = "product"):
    """variant "product": the schedule the kernels ship with (NAF Miller loop);
    variant "refchain": -DBN_ATE_NAF=0, the reference's binary walk, whose 102 lines and unreduced Miller value can be
    compared with the reference's known answers."""
    if variant not in _libs:
        so = os.path.join(_HERE, "libemu.so" if variant == "product" else "libemu_refchain.so")
        src = os.path.join(_HERE, "emu.cpp")
        csrc = os.path.join(os.path.dirname(_HERE), "..", "bn_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".inc"))]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            flags = os.environ.get("BN_EMU_FLAGS", "").split() + ([] if variant == "product" else ["-DBN_ATE_NAF=0"])
            # -frounding-math / -mfma: f52.cuh's round-toward-zero fma runs under fesetround(FE_TOWARDZERO) on the host
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-frounding-math", "-mfma", *flags, "-shared", "-fPIC", "-o", so, src, "-lpthread"])
        _libs[variant] = ctypes.CDLL(so)
    return _libs[variant]


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint64)


def fp_op(op, which, a, b=None):
    a, b = _c(a), _c(b)
    out = np.zeros(4, dtype=np.uint64)
    lib().emu_fp_op(op, which, _p(a), _p(b), _p(out))
    return out


def f52_mul(a, b):
    """a * b * 2^-256 mod q through the FP64 product path (f52.cuh)."""
    a, b = _c(a), _c(b)
    out = np.zeros(4, dtype=np.uint64)
    lib().emu_f52_mul(_p(a), _p(b), _p(out))
    return out


def fp2_op(op, a, b=None):
    a, b = _c(a), _c(b)
    out = np.zeros(8, dtype=np.uint64)
    lib().emu_fp2_op(op, _p(a), _p(b), _p(out))
    return out


def g1_mul(p, fr):
    p, fr = _c(p), _c(fr)
    out = np.zeros(12, dtype=np.uint64)
    lib().emu_g1_mul(_p(p), _p(fr), _p(out))
    return out


def g2_mul(p, fr):
    p, fr = _c(p), _c(fr)
    out = np.zeros(24, dtype=np.uint64)
    lib().emu_g2_mul(_p(p), _p(fr), _p(out))
    return out


def fp_mul2(a, b, c, d):
    out = np.zeros(4, dtype=np.uint64)
    lib().emu_fp_mul2(_p(_c(a)), _p(_c(b)), _p(_c(c)), _p(_c(d)), _p(out))
    return out


def lines(g1, g2, variant="product"):
    g1, g2 = _c(g1), _c(g2)
    L = lib(variant)
    out = np.zeros((L.emu_num_lines(), 40), dtype=np.uint64)
    pa = np.zeros(8, dtype=np.uint64)
    qa = np.zeros(16, dtype=np.uint64)
    finite = L.emu_lines(_p(g1), _p(g2), _p(out), _p(pa), _p(qa))
    return finite, out, pa, qa


def lines_duo(g1, g2, variant="product"):
    g1, g2 = _c(g1), _c(g2)
    L = lib(variant)
    out = np.zeros((L.emu_num_lines(), 40), dtype=np.uint64)
    finite = L.emu_lines_duo(_p(g1), _p(g2), _p(out))
    return finite, out


def gt_op(op, a, b=None, arg=0):
    a, b = _c(a), _c(b)
    out = np.zeros(48, dtype=np.uint64)
    lib().emu_gt_op(op, _p(a), _p(b), arg, _p(out))
    return out


def miller(lines_arr, variant="product"):
    l = _c(lines_arr)
    out = np.zeros(48, dtype=np.uint64)
    lib(variant).emu_miller(_p(l), _p(out))
    return out


def pairing(g1, g2, variant="product"):
    g1, g2 = _c(g1), _c(g2)
    out = np.zeros(48, dtype=np.uint64)
    lib(variant).emu_pairing(_p(g1), _p(g2), _p(out))
    return out


_WIRE = {"fr": (0, 4, 32), "g1": (1, 12, 65), "g2": (2, 24, 129)}


def wire_encode(kind, img):
    code, words, nbytes = _WIRE[kind]
    img = _c(img)
    rec = np.zeros(nbytes, dtype=np.uint8)
    lib().emu_wire_encode(code, _p(img), rec.ctypes.data_as(ctypes.c_void_p))
    return rec


def wire_decode(kind, rec):
    code, words, nbytes = _WIRE[kind]
    rec = np.ascontiguousarray(rec, dtype=np.uint8)
    assert rec.shape == (nbytes,)
    img = np.zeros(words, dtype=np.uint64)
    st = lib().emu_wire_decode(code, rec.ctypes.data_as(ctypes.c_void_p), _p(img))
    return img, st
