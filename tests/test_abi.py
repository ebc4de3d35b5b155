"""CPU tests: the C-ABI library is built, loads, and exports every symbol include/bn_b200.h declares;
without a GPU every entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bn_b200 import _lib, build
    if not os.path.exists(_lib.SO_PATH):
        build.build()
    return _lib.load()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "bn_b200.h")) as f:
        hdr = f.read()
    return sorted(set(re.findall(r"\b(bn_b200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    from bn_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert sorted(_lib.EXPORTS) == syms


def test_struct_sizes_match_reference_layouts():
    with open(os.path.join(ROOT, "include", "bn_b200.h")) as f:
        hdr = f.read()
    # sizes implied by the typedefs: Fr 32, G1 96, G2 192, Gt 384
    assert "uint64_t l[4]" in hdr and "x[4], y[4], z[4]" in hdr and "x[2][4], y[2][4], z[2][4]" in hdr and "c[2][3][2][4]" in hdr


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bn_b200
    rc = lib.bn_b200_init(0)
    assert rc == -1
    assert b"no CPU fallback" in lib.bn_b200_last_error()
    g1 = np.zeros((1, 12), dtype=np.uint64)
    g2 = np.zeros((1, 24), dtype=np.uint64)
    with pytest.raises(bn_b200.BnB200Error):
        bn_b200.pairing_batch(g1, g2)
    out = np.zeros((1, 48), dtype=np.uint64)
    rc = lib.bn_b200_pairing_batch(g1.ctypes.data_as(ctypes.c_void_p), g2.ctypes.data_as(ctypes.c_void_p),
                                   out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(1))
    assert rc == -1 and not out.any()


def test_product_does_not_reference_oracle():
    """The shipped package must not import / include / dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "bn_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|#include\s*[<\"][^>\"]*oracle|libbn_ref|bn_ref\.|cref)")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                with open(os.path.join(dirpath, fn)) as f:
                    assert not bad.search(f.read()), fn
