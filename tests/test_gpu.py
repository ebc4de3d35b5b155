"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI
(libbn_b200.so via bn_b200/_lib.py) and is compared bit-for-bit with the oracle on the same inputs.
Mirrors the reference's tests: test_reduced_pairing / test_binlinearity (src/groups/mod.rs:773-823),
group_trials (src/groups/tests.rs), golden vectors (tests/serialization.rs)."""
import ctypes

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


def test_fq_mul_chain(bn):
    a = util.synth_scalars(0xB2000002, 300)
    b = util.synth_scalars(0xB2000012, 300)
    a[0] = 0
    b[1] = util.words((o.Q - 1).to_bytes(32, "little"))
    for iters in (1, 7, 100):
        assert np.array_equal(bn.fq_mul_chain(a, b, iters), cref.fq_mul_chain(a, b, iters, 4)), iters


def test_pairing_kat(bn):
    kat = util.load_json("pairing_kat.json")
    g1 = cref.g1_mul_batch(cref.g1_generator(), util.fr_img(int(kat["k1"]))[None])
    g2 = cref.g2_mul_batch(cref.g2_generator(), util.fr_img(int(kat["k2"]))[None])
    gt = bn.pairing_batch(g1, g2)
    assert np.array_equal(gt[0], util.gt_img(o.fq12_from_flat(kat["reduced_pairing"])))
    # value-class mirror of the crate API
    assert bn.pairing(bn.G1(g1[0]), bn.G2(g2[0])) == bn.Gt(gt[0])


def test_pairing_edge_cases(bn):
    g1, g2 = util.edge_case_pairs()
    assert np.array_equal(bn.pairing_batch(g1, g2), cref.pairing_batch(g1, g2))
    one = util.gt_img(o.FQ12_ONE)
    got = bn.pairing_batch(g1, g2)
    for i in (1, 2, 4, 5):
        assert np.array_equal(got[i], one)


@pytest.mark.parametrize("n", [1, 4, 5, 6, 19, 20, 21, 257])
def test_pairing_ragged_batches(bn, n):
    g1, g2 = util.synth_pairs(0xB2000001 + n, n)
    assert np.array_equal(bn.pairing_batch(g1, g2), cref.pairing_batch(g1, g2, 8))


def test_pairing_config1_1024(bn):
    """BASELINE config 1: 1024 random pairs, Gt bit-exact vs the CPU oracle (with edge cases mixed in)."""
    g1, g2 = util.synth_pairs(0xB2000001, 1024)
    e1, e2 = util.edge_case_pairs()
    g1[100:100 + len(e1)] = e1
    g2[100:100 + len(e2)] = e2
    assert np.array_equal(bn.pairing_batch(g1, g2), cref.pairing_batch(g1, g2, 8))
    assert bn.pairing_batch(np.zeros((0, 12), np.uint64), np.zeros((0, 24), np.uint64)).shape == (0, 48)


def test_scalar_mul_limb_exact(bn):
    n = 200
    g1, g2 = util.synth_pairs(0xB2000003, n)
    k = util.synth_scalars(0xB2000013, n)
    for i, s in enumerate([0, 1, 2, o.R_ORDER - 1, 23938123]):
        k[i] = util.fr_img(s)
    g1[10] = util.g1_img(o.g_zero(o.FQ))
    g2[11] = util.g2_img(o.g_zero(o.FQ2))
    g1[12] = cref.g1_generator()[0]
    g2[12] = cref.g2_generator()[0]
    assert np.array_equal(bn.g1_mul_batch(g1, k), cref.g1_mul_batch(g1, k, 8))
    assert np.array_equal(bn.g2_mul_batch(g2, k), cref.g2_mul_batch(g2, k, 8))


def test_g1_golden_vectors(bn):
    """tests/serialization.rs g1_vectors recurrence acc <- acc*23938123 + acc; the GPU does the scalar mul."""
    lines = util.load_vectors("g1_vectors.txt", 60)
    k = util.fr_img(23938123)[None]
    acc = cref.g1_generator()
    for want in lines:
        assert o.encode_g1(util.img_g1(cref.g1_normalize(acc)[0])).hex() == want
        acc = cref.g1_add(bn.g1_mul_batch(acc, k), acc)


def test_g2_golden_vectors(bn):
    lines = util.load_vectors("g2_vectors.txt", 25)
    k = util.fr_img(23938123)[None]
    acc = cref.g2_generator()
    for want in lines:
        assert o.encode_g2(util.img_g2(cref.g2_normalize(acc)[0])).hex() == want
        acc = cref.g2_add(bn.g2_mul_batch(acc, k), acc)


def test_fr_golden_vectors_on_gpu(bn):
    """tests/serialization.rs fr_vectors: acc <- acc*acc + acc + acc^-1, every step computed by the GPU Fr kernels."""
    lines = util.load_vectors("fr_vectors.txt", 400)
    acc = util.fr_img(1)[None]
    for want in lines:
        assert o.encode_fr(o.fr_from_bytes(acc[0].tobytes())).hex() == want
        sq = bn.fr_op_batch("mul", acc, acc)
        acc = bn.fr_op_batch("add", bn.fr_op_batch("add", sq, acc), bn.fr_op_batch("inverse", acc))
    k = util.synth_scalars(123, 300)
    k[0] = 0
    inv = bn.fr_op_batch("inverse", k)
    assert not inv[0].any()
    one = util.fr_img(1)
    assert all(np.array_equal(x, one) for x in bn.fr_op_batch("mul", k[1:], inv[1:]))
    assert np.array_equal(bn.fr_op_batch("sub", k, k), np.zeros_like(k))
    assert np.array_equal(bn.fr_op_batch("add", k, bn.fr_op_batch("neg", k)), np.zeros_like(k))


def test_normalize_and_wire_vectors(bn):
    """Group::normalize on the GPU reproduces the reference's 65/129-byte encodings of the golden recurrences."""
    k = util.fr_img(23938123)[None]
    for name, gen, mulb, addc, norm_gpu, norm_cpu, enc, img, nvec in (
            ("g1", cref.g1_generator(), bn.g1_mul_batch, cref.g1_add, bn.g1_normalize_batch, cref.g1_normalize, o.encode_g1, util.img_g1, 40),
            ("g2", cref.g2_generator(), bn.g2_mul_batch, cref.g2_add, bn.g2_normalize_batch, cref.g2_normalize, o.encode_g2, util.img_g2, 15)):
        lines = util.load_vectors(name + "_vectors.txt", nvec)
        acc, accs = gen, []
        for _ in lines:
            accs.append(acc[0])
            acc = addc(mulb(acc, k), acc)
        accs = np.stack(accs)
        normed = norm_gpu(accs)
        assert np.array_equal(normed, np.concatenate([norm_cpu(a[None]) for a in accs]))
        for row, want in zip(normed, lines):
            assert enc(img(row)).hex() == want
    inf1 = util.g1_img(o.g_zero(o.FQ))[None]
    assert np.array_equal(bn.g1_normalize_batch(inf1), inf1)
    inf2 = util.g2_img(o.g_zero(o.FQ2))[None]
    assert np.array_equal(bn.g2_normalize_batch(inf2), inf2)


def _fq2_sqrt(a):
    """Square root in Fq2 for q = 3 mod 4 (None if a is a non-residue); test helper for off-subgroup G2 points."""
    if a == (0, 0):
        return a
    a1 = o.fq2_pow(a, (o.Q - 3) // 4)
    alpha = o.fq2_mul(o.fq2_sqr(a1), a)
    a0 = o.fq2_mul(o.fq2_conj(alpha), alpha)
    if a0 == (o.Q - 1, 0):
        return None
    x0 = o.fq2_mul(a1, a)
    if alpha == (o.Q - 1, 0):
        return o.fq2_mul((0, 1), x0)
    b = o.fq2_pow(o.fq2_add(o.FQ2_ONE, alpha), (o.Q - 1) // 2)
    return o.fq2_mul(b, x0)


def test_decode_checks(bn):
    """On-curve / subgroup checks of AffineG::decode (src/groups/mod.rs:178-205) on the GPU, incl. the reference's
    'not on the curve' edge vectors (tests/serialization.rs:66-68) and G2 points on the twist but outside the subgroup."""
    n = 40
    g1, g2 = util.synth_pairs(0xDEC0, n)
    a1 = bn.g1_normalize_batch(g1)
    a2 = bn.g2_normalize_batch(g2)
    a1[3] = util.g1_img(o.g_zero(o.FQ))
    a2[4] = util.g2_img(o.g_zero(o.FQ2))
    bad1 = a1.copy()
    bad1[:, 4] ^= np.uint64(2)  # perturb y
    bad2 = a2.copy()
    bad2[:, 8] ^= np.uint64(2)
    assert bn.g1_check_batch(a1).all() and bn.g2_check_batch(a2).all()
    r1, r2 = bn.g1_check_batch(bad1), bn.g2_check_batch(bad2)
    assert not r1[np.arange(n) != 3].any() and r1[3]   # the zero point stays acceptable
    assert not r2[np.arange(n) != 4].any() and r2[4]
    # the reference's own off-curve vectors (x, y taken from the hex, bypassing the host-side decoder)
    for kind, hx in util.load_json("wire_edge_cases.json")["cases"]:
        b = bytes.fromhex(hx)
        if kind == "G1" and len(b) == 65:
            x, y = int.from_bytes(b[1:33], "big"), int.from_bytes(b[33:65], "big")
            assert not bn.g1_check_batch(util.g1_img((x, y, 1))[None])[0]
    # points on the twist y^2 = x^3 + 3/xi that are NOT in the order-r subgroup
    found = 0
    for xv in range(1, 60):
        x = (xv, 1)
        y = _fq2_sqrt(o.fq2_add(o.fq2_mul(o.fq2_sqr(x), x), o.G2_B))
        if y is None:
            continue
        p = (x, y, o.FQ2_ONE)
        assert o.fq2_sqr(y) == o.fq2_add(o.fq2_mul(o.fq2_sqr(x), x), o.G2_B)
        in_subgroup = o.g_is_zero(o.FQ2, o.g_add(o.FQ2, o.g_mul(o.FQ2, p, o.R_ORDER - 1), p))
        assert bool(bn.g2_check_batch(util.g2_img(p)[None])[0]) == in_subgroup
        found += not in_subgroup
        if found >= 3:
            break
    assert found >= 3


def test_gt_mul_pow(bn):
    n = 23
    g1, g2 = util.synth_pairs(77, n)
    gt = cref.pairing_batch(g1, g2, 8)
    k = util.synth_scalars(78, n)
    for i, s in enumerate([0, 1, 2, o.R_ORDER - 1]):
        k[i] = util.fr_img(s)
    assert np.array_equal(bn.gt_mul_batch(gt, gt[::-1].copy()), cref.gt_mul_batch(gt, gt[::-1].copy(), 8))
    assert np.array_equal(bn.gt_pow_batch(gt, k), cref.gt_pow_batch(gt, k, 8))
    assert np.array_equal(bn.gt_inv_batch(gt), np.concatenate([cref.fq12_inv(x[None]) for x in gt]))
    one = util.gt_img(o.FQ12_ONE)
    assert (bn.gt_mul_batch(gt, bn.gt_inv_batch(gt)) == one[None]).all()
    # non-cyclotomic operand
    f = util.load_json("fq12_kat.json")
    s = util.gt_img(o.fq12_from_flat(f["vector_start"]))[None]
    assert np.array_equal(bn.gt_mul_batch(s, s), cref.fq12_mul(s, s))
    assert np.array_equal(bn.gt_pow_batch(s, k[5:6]), cref.gt_pow_batch(s, k[5:6]))
    assert np.array_equal(bn.gt_inv_batch(s), cref.fq12_inv(s))


def test_bilinearity_on_gpu(bn):
    """test_binlinearity (src/groups/mod.rs:798-823) entirely on the GPU: e(P,Q)^s == e(sP,Q) == e(P,sQ)."""
    n = 64
    g1, g2 = util.synth_pairs(5, n)
    s = util.synth_scalars(6, n)
    a = bn.gt_pow_batch(bn.pairing_batch(g1, g2), s)
    b = bn.pairing_batch(bn.g1_mul_batch(g1, s), g2)
    c = bn.pairing_batch(g1, bn.g2_mul_batch(g2, s))
    assert np.array_equal(a, b) and np.array_equal(b, c)
    one = util.gt_img(o.FQ12_ONE)
    m1 = np.repeat(util.fr_img(o.R_ORDER - 1)[None], n, axis=0)
    assert all(np.array_equal(x, one) for x in bn.gt_mul_batch(bn.gt_pow_batch(a, m1), a))
    assert not any(np.array_equal(x, one) for x in a)


def test_fused_pairing_pow(bn):
    """Joux pattern (reference examples/joux.rs:19-21): pairing(P, Q).pow(s) fused == pairing then Gt::pow == oracle."""
    n = 27
    g1, g2 = util.synth_pairs(0x10AD, n)
    e1, e2 = util.edge_case_pairs()
    g1[5:5 + len(e1)], g2[5:5 + len(e2)] = e1, e2
    s = util.synth_scalars(0x10AE, n)
    for i, v in enumerate([0, 1, 2, 3, o.R_ORDER - 1]):
        s[i] = util.fr_img(v)
    fused = bn.pairing_pow_batch(g1, g2, s)
    assert np.array_equal(fused, bn.gt_pow_batch(bn.pairing_batch(g1, g2), s))
    assert np.array_equal(fused, cref.gt_pow_batch(cref.pairing_batch(g1, g2, 8), s, 8))
    # three-party key agreement: e(bP, cQ)^a == e(cP, aQ)^b == e(aP, bQ)^c
    sk = util.synth_scalars(0x10AF, 3)
    gen1, gen2 = cref.g1_generator(), cref.g2_generator()
    pk1 = bn.g1_mul_batch(np.repeat(gen1, 3, axis=0), sk)
    pk2 = bn.g2_mul_batch(np.repeat(gen2, 3, axis=0), sk)
    ss = bn.pairing_pow_batch(pk1[[1, 2, 0]], pk2[[2, 0, 1]], sk)
    assert np.array_equal(ss[0], ss[1]) and np.array_equal(ss[1], ss[2])


# (parity at the BASELINE config sizes: tests/test_gpu_sizes.py)


def test_device_pointer_api_with_torch(bn):
    import torch
    n = 40
    g1, g2 = util.synth_pairs(9, n)
    t1 = torch.from_numpy(g1.view(np.int64)).cuda()
    t2 = torch.from_numpy(g2.view(np.int64)).cuda()
    out = torch.empty((n, 48), dtype=torch.int64, device="cuda")
    lib = bn.load()
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.bn_b200_pairing_batch_dev(ctypes.c_void_p(t1.data_ptr()), ctypes.c_void_p(t2.data_ptr()),
                                       ctypes.c_void_p(out.data_ptr()), ctypes.c_size_t(n), ctypes.c_void_p(st))
    assert rc == 0
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), cref.pairing_batch(g1, g2, 8))
    assert lib.bn_b200_launch_count() >= 2


@pytest.mark.gpu
def test_fused_gather_entry_point_single_rank(bn):
    """bn_b200_pairing_batch_gather_dev with world = 1 (the peer table holds only this device's buffer): same bytes as
    pairing_batch.  The multi-rank form is exercised by bench.py --gpus N (asserts equality with the NCCL all_gather)."""
    import ctypes
    import torch
    lib = bn.load()
    g1, g2 = util.synth_pairs(0xB200000A, 37)
    want = bn.pairing_batch(g1, g2)
    dev = torch.device("cuda", 0)
    d1 = torch.from_numpy(g1.view(np.int64)).to(dev)
    d2 = torch.from_numpy(g2.view(np.int64)).to(dev)
    out = torch.zeros((len(g1), 48), dtype=torch.int64, device=dev)
    ptrs = (ctypes.c_void_p * 1)(out.data_ptr())
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        rc = lib.bn_b200_pairing_batch_gather_dev(ctypes.c_void_p(d1.data_ptr()), ctypes.c_void_p(d2.data_ptr()), ptrs, 1, 0,
                                                  ctypes.c_size_t(len(g1)), ctypes.c_void_p(st.cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
    assert lib.bn_b200_pairing_batch_gather_dev(None, None, ptrs, 9, 0, ctypes.c_size_t(1), None) != 0  # world > 8 rejected


@pytest.mark.gpu
def test_pairing_batch_pinned_output_zero_copy(bn):
    """Host-pointer pairing_batch with page-locked buffers: the last kernel writes the pinned output directly (no D2H
    pass); same bytes as the staged path used for pageable memory, also with BN_B200_ZEROCOPY semantics off (numpy)."""
    import ctypes
    import torch
    lib = bn.load()
    g1, g2 = util.synth_pairs(0xB200000B, 53)
    e1, e2 = util.edge_case_pairs()
    g1, g2 = np.concatenate([g1, e1]), np.concatenate([g2, e2])
    want = bn.pairing_batch(g1, g2)  # pageable numpy buffers: staged path
    h1 = torch.from_numpy(g1.view(np.int64).copy()).pin_memory()
    h2 = torch.from_numpy(g2.view(np.int64).copy()).pin_memory()
    ho = torch.zeros((len(g1), 48), dtype=torch.int64).pin_memory()
    rc = lib.bn_b200_pairing_batch(ctypes.c_void_p(h1.data_ptr()), ctypes.c_void_p(h2.data_ptr()), ctypes.c_void_p(ho.data_ptr()),
                                   ctypes.c_size_t(len(g1)))
    assert rc == 0
    assert np.array_equal(ho.numpy().view(np.uint64), want)


@pytest.mark.gpu
def test_gt_mul_pow_extreme_coefficients(bn):
    """The kernels' lazy arithmetic at its bounds: Gt images whose coefficients are all q-1 / 0 / mixed (not pairing
    values: Gt mul and pow are plain Fq12 operations, reference src/lib.rs:171-179) against the big-int oracle."""
    import random
    rng = random.Random(6)
    qm1 = o.Q - 1
    pats = [[qm1] * 12, [0] * 12, [1] + [0] * 11, [qm1, 0] * 6, [0, qm1] * 6, [qm1] * 6 + [0] * 6,
            [rng.choice((0, 1, qm1, qm1 - 1, o.Q // 2)) for _ in range(12)], [rng.randrange(o.Q) for _ in range(12)]]
    elems = [o.fq12_from_flat(p) for p in pats]
    a = np.stack([util.gt_img(x) for x in elems])
    b = np.roll(a, 3, axis=0)
    got = bn.gt_mul_batch(a, b)
    for i in range(len(elems)):
        assert np.array_equal(got[i], util.gt_img(o.fq12_mul(elems[i], elems[(i - 3) % len(elems)]))), i
    k = np.stack([util.fr_img(5)] * len(elems))
    got = bn.gt_pow_batch(a, k)
    for i in range(len(elems)):
        assert np.array_equal(got[i], util.gt_img(o.fq12_pow(elems[i], 5))), i


@pytest.mark.gpu
def test_pairing_chunked_batches(bn):
    """Batches above the chunk size are processed in several passes over the bounded line buffer: identical results,
    through the device-resident, the staged host and the pinned zero-copy paths."""
    import ctypes
    import torch
    lib = bn.load()
    g1, g2 = util.synth_pairs(0xB200000C, 64)
    g1, g2 = np.tile(g1, (4, 1))[:237], np.tile(g2, (4, 1))[:237]
    want = bn.pairing_batch(g1, g2)
    assert lib.bn_b200_set_max_chunk(ctypes.c_size_t(100)) == 0
    try:
        assert np.array_equal(bn.pairing_batch(g1, g2), want)          # 100 + 100 + 37
        h1 = torch.from_numpy(g1.view(np.int64).copy()).pin_memory()
        h2 = torch.from_numpy(g2.view(np.int64).copy()).pin_memory()
        ho = torch.zeros((len(g1), 48), dtype=torch.int64).pin_memory()
        assert lib.bn_b200_pairing_batch(ctypes.c_void_p(h1.data_ptr()), ctypes.c_void_p(h2.data_ptr()),
                                         ctypes.c_void_p(ho.data_ptr()), ctypes.c_size_t(len(g1))) == 0
        assert np.array_equal(ho.numpy().view(np.uint64), want)
        k = util.synth_scalars(0xB200000D, len(g1))
        assert np.array_equal(bn.pairing_pow_batch(g1, g2, k), bn.gt_pow_batch(want, k))
    finally:
        lib.bn_b200_set_max_chunk(ctypes.c_size_t(0))


@pytest.mark.gpu
@pytest.mark.parametrize("parts", [2, 3, 4])
def test_pairing_split_sub_batches(bn, parts):
    """A call run as several sub-batches on the library's side streams (automatic at the headline
    batch) gives the results of the single sequence of kernels: in place, staged host, pinned zero-copy and fused pow,
    with ragged sub-batches (237 = 80 + 80 + 77, 120 + 117, 60 + 60 + 60 + 57) and with chunking on top."""
    import ctypes
    import torch
    lib = bn.load()
    g1, g2 = util.synth_pairs(0xB200000E, 64)
    g1, g2 = np.tile(g1, (4, 1))[:237], np.tile(g2, (4, 1))[:237]
    e1, e2 = util.edge_case_pairs()
    g1[100:100 + len(e1)], g2[100:100 + len(e2)] = e1, e2
    assert lib.bn_b200_set_split(1, ctypes.c_size_t(0)) == 0
    try:
        want = bn.pairing_batch(g1, g2)
        assert np.array_equal(want[:64], cref.pairing_batch(g1[:64], g2[:64], 8))
        assert lib.bn_b200_set_split(parts, ctypes.c_size_t(20)) == 0
        assert np.array_equal(bn.pairing_batch(g1, g2), want)
        h1 = torch.from_numpy(g1.view(np.int64).copy()).pin_memory()
        h2 = torch.from_numpy(g2.view(np.int64).copy()).pin_memory()
        ho = torch.zeros((len(g1), 48), dtype=torch.int64).pin_memory()
        assert lib.bn_b200_pairing_batch(ctypes.c_void_p(h1.data_ptr()), ctypes.c_void_p(h2.data_ptr()),
                                         ctypes.c_void_p(ho.data_ptr()), ctypes.c_size_t(len(g1))) == 0
        assert np.array_equal(ho.numpy().view(np.uint64), want)
        k = util.synth_scalars(0xB200000F, len(g1))
        assert np.array_equal(bn.pairing_pow_batch(g1, g2, k), bn.gt_pow_batch(want, k))
        assert lib.bn_b200_set_max_chunk(ctypes.c_size_t(150)) == 0      # 150 (split) + 87 (split)
        assert np.array_equal(bn.pairing_batch(g1, g2), want)
        assert lib.bn_b200_set_split(5, ctypes.c_size_t(0)) != 0
    finally:
        lib.bn_b200_set_max_chunk(ctypes.c_size_t(0))
        lib.bn_b200_set_split(0, ctypes.c_size_t(0))
