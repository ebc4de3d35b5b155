"""The C++ host layer (include/bn.hpp) over the C ABI: builds and links on CPU; on the GPU box it replays the
reference's usage pattern (pairing, G*Fr, Gt::pow) on oracle-generated vectors and must be bit-exact."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "api_main.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "api_main")


def build_binary():
    from bn_b200 import _lib, build
    if not os.path.exists(_lib.SO_PATH):
        build.build()
    libdir = os.path.dirname(_lib.SO_PATH)
    if (not os.path.exists(BIN)) or os.path.getmtime(BIN) < max(os.path.getmtime(SRC), os.path.getmtime(_lib.SO_PATH)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", BIN, SRC, "-L" + libdir, "-lbn_b200",
                               "-Wl,-rpath," + libdir])
    return BIN


def test_cpp_api_builds_and_links():
    exe = build_binary()
    out = subprocess.check_output([exe, "/dev/null", "--link-only"], text=True)
    assert "link ok" in out


@pytest.mark.gpu
def test_cpp_api_bit_exact(tmp_path):
    from oracle import cref
    from tests import util
    exe = build_binary()
    n = 12
    g1, g2 = util.synth_pairs(0xC0FFEE, n)
    e1, e2 = util.edge_case_pairs()
    g1[8:12], g2[8:12] = e1[:4], e2[:4]
    fr = util.synth_scalars(0xC0FFEF, n)
    gt = cref.pairing_batch(g1, g2, 4)
    sg1 = cref.g1_mul_batch(g1, fr, 4)
    pw = cref.gt_pow_batch(gt, fr, 4)
    path = tmp_path / "vec.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", n))
        for a in (g1, g2, fr, gt, sg1, pw):
            f.write(np.ascontiguousarray(a, dtype="<u8").tobytes())
    res = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "cpp api ok" in res.stdout
