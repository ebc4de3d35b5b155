"""Parity at the BASELINE config sizes: every element of configs 2-4 is compared with the C
oracle (all host threads: 2^14 pairings take a few seconds), config 5's 2^17 on >= 1024 samples plus whole-batch
properties.  Inputs are generated on the device by the product's own scalar-mul kernels and verified as part of the test."""
import os

import numpy as np
import pytest

from oracle import bn_oracle as o
from oracle import cref
from tests import util

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 8


@pytest.fixture(scope="module")
def bn():
    import bn_b200
    bn_b200.init(0)
    return bn_b200


def tiled_scalars(seed, n, base=1 << 11):
    """n canonical Fr images: a tile of `base` big-int generated scalars, low limb perturbed per element (still < r)."""
    k = np.tile(util.synth_scalars(seed, min(base, n)), ((n + base - 1) // base, 1))[:n].copy()
    k[:, 0] ^= (np.arange(n, dtype=np.uint64) >> np.uint64(11))
    return k


def test_config2_fq_mul_chain_full_size(bn):
    """BASELINE config 2: 2^20 lanes.  All 2^20 outputs of a 64-step chain against the oracle (x <- x*b)."""
    n = 1 << 20
    a, b = tiled_scalars(0xB2000002, n), tiled_scalars(0xB2000012, n)  # any value < r < q is a canonical Fq image too
    got = bn.fq_mul_chain(a, b, 64)
    assert np.array_equal(got, cref.fq_mul_chain(a, b, 64, THREADS))
    # the dedicated squaring chain: x^(2^k) -- against repeated multiplication on the oracle
    m = 1 << 14
    sq = bn.fq_sqr_chain(a[:m], 7)
    want = a[:m].copy()
    for _ in range(7):
        want = cref.fq_mul_chain(want, want, 1, THREADS)
    assert np.array_equal(sq, want)


def test_config3_g1_scalar_mul_full_size(bn):
    """BASELINE config 3: 2^16 G1 scalar multiplications, every Jacobian limb against the oracle."""
    n = 1 << 16
    k = tiled_scalars(0xB2000003, n)
    base = np.repeat(cref.g1_generator(), n, axis=0)
    p = bn.g1_mul_batch(base, tiled_scalars(0xB2000013, n))        # random Jacobian points (z != 1)
    assert np.array_equal(p[:: n // 512], cref.g1_mul_batch(base[:: n // 512], tiled_scalars(0xB2000013, n)[:: n // 512], THREADS))
    got = bn.g1_mul_batch(p, k)
    assert np.array_equal(got, cref.g1_mul_batch(p, k, THREADS))


def test_config4_all_16384_pairings_vs_oracle(bn):
    """BASELINE config 4: 2^14 pairings, ALL compared with the oracle bit for bit."""
    n = 1 << 14
    a, b = tiled_scalars(0xB2000004, n), tiled_scalars(0xB2000014, n)
    g1 = bn.g1_mul_batch(np.repeat(cref.g1_generator(), n, axis=0), a)   # on-device input generation (row f-2)
    g2 = bn.g2_mul_batch(np.repeat(cref.g2_generator(), n, axis=0), b)
    e1, e2 = util.edge_case_pairs()
    g1[5000:5000 + len(e1)], g2[5000:5000 + len(e2)] = e1, e2
    gt = bn.pairing_batch(g1, g2)
    assert np.array_equal(gt, cref.pairing_batch(g1, g2, THREADS))


def test_config5_size_sampled_1024_plus_properties(bn):
    """BASELINE config 5's 2^17 pairs on ONE GPU: 1024 + tail samples against the oracle, and over the WHOLE batch
    e(P, Q) * e(-P, Q) == 1 with the negation done by the device group law."""
    n = 1 << 17
    a, b = tiled_scalars(0xB2000005, n), tiled_scalars(0xB2000015, n)
    g1 = bn.g1_mul_batch(np.repeat(cref.g1_generator(), n, axis=0), a)
    g2 = bn.g2_mul_batch(np.repeat(cref.g2_generator(), n, axis=0), b)
    gt = bn.pairing_batch(g1, g2)
    idx = np.concatenate([np.arange(0, n, n // 1024), [n - 1, n - 2, n - 5, n - 6, n - 19, n - 20, n - 21]])
    assert np.array_equal(gt[idx], cref.pairing_batch(g1[idx], g2[idx], THREADS))
    prod = bn.gt_mul_batch(gt, bn.pairing_batch(bn.g1_op_batch("neg", g1), g2))
    assert (prod == util.gt_img(o.FQ12_ONE)[None]).all()
    assert len(np.unique(gt[:4096], axis=0)) == 4096   # the results are not one repeated value
