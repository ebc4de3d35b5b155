"""Exactness of the FP64 multiply-accumulate prototype (tools/proto/fp64_mac.cuh): products of
256-bit operands through 44-bit double limbs and FMA high/low splitting, accumulated over several rounds, must equal the
big-integer result bit for bit -- including all-ones operands and the accumulation depth of a dense Fq12 product."""
import ctypes
import os
import random
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "tools", "proto", "fp64_mac_host.cpp")
SO = os.path.join(HERE, "..", "tools", "proto", "libfp64proto.so")


def _lib():
    dep = [SRC, os.path.join(os.path.dirname(SRC), "fp64_mac.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in dep):
        # -ffp-contract=off: the splitting relies on the written sequence of FMAs and additions
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC])
    return ctypes.CDLL(SO)


def _limbs(x, n=8):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def _run(pairs):
    lib = _lib()
    a = np.array([l for x, _ in pairs for l in _limbs(x)], dtype=np.uint32)
    b = np.array([l for _, y in pairs for l in _limbs(y)], dtype=np.uint32)
    out = np.zeros(17, dtype=np.uint32)
    lib.fp64_mac_rounds(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), len(pairs),
                        out.ctypes.data_as(ctypes.c_void_p))
    return sum(int(v) << (32 * i) for i, v in enumerate(out))


def test_fp64_mac_exact():
    rng = random.Random(44)
    top = (1 << 256) - 1
    cases = [[(top, top)], [(top, top)] * 6, [(0, top), (top, 0), (1, 1)], [(1 << 255, 1 << 255)] * 6,
             [((1 << 44) - 1, (1 << 44) - 1)], [(top, 1), (1, top)]]
    for _ in range(40):
        cases.append([(rng.getrandbits(256), rng.getrandbits(256)) for _ in range(rng.choice((1, 3, 4, 6)))])
    for _ in range(10):  # limbs at their extremes
        pick = lambda: sum(rng.choice((0, (1 << 44) - 1, 1 << 43, 1)) << (44 * k) for k in range(6)) & top
        cases.append([(pick(), pick()) for _ in range(6)])
    for pairs in cases:
        want = sum(x * y for x, y in pairs) % (1 << 544)
        assert _run(pairs) == want, pairs
