"""The C++ host layer (include/bn.hpp) over the C ABI: builds and links on CPU; on the GPU box it replays the
reference's usage pattern (pairing, G*Fr, Gt::pow, the group law and Fr operators) on oracle-generated vectors and must
be bit-exact; with more than one GPU it drives all of them from ONE process (bn_b200_init_multi)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "api_main.cpp")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    """Always rebuilt against the current header and library (never a stale prebuilt binary)."""
    from bn_b200 import _lib, build
    if not os.path.exists(_lib.SO_PATH):
        build.build()
    libdir = os.path.dirname(_lib.SO_PATH)
    out = str(tmp_path_factory.mktemp("cpp") / "api_main")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", out, SRC, "-L" + libdir, "-lbn_b200", "-Wl,-rpath," + libdir])
    return out


def test_cpp_api_builds_and_links(exe):
    out = subprocess.check_output([exe, "/dev/null", "--link-only"], text=True)
    assert "link ok" in out


def fr_pow_oracle(a_img, e_img):
    """Fr::pow on Montgomery images with plain integers (reference src/fields/mod.rs:35-46 via src/lib.rs:24)."""
    from oracle import bn_oracle as o
    from tests import util
    rinv = pow(1 << 256, -1, o.R_ORDER)
    out = []
    for a, e in zip(a_img, e_img):
        av = int.from_bytes(a.tobytes(), "little") * rinv % o.R_ORDER
        ev = int.from_bytes(e.tobytes(), "little") * rinv % o.R_ORDER
        out.append(util.fr_img(pow(av, ev, o.R_ORDER)))
    return np.stack(out)


@pytest.mark.gpu
def test_cpp_api_bit_exact(exe, tmp_path):
    from oracle import cref
    from tests import util
    n = 12
    g1, g2 = util.synth_pairs(0xC0FFEE, n)
    e1, e2 = util.edge_case_pairs()
    g1[8:12], g2[8:12] = e1[:4], e2[:4]
    fr = util.synth_scalars(0xC0FFEF, n)
    gt = cref.pairing_batch(g1, g2, 4)
    sg1 = cref.g1_mul_batch(g1, fr, 4)
    pw = cref.gt_pow_batch(gt, fr, 4)
    add1 = cref.g1_add(g1, np.roll(g1, -1, axis=0))
    sub2 = cref.g2_add(g2, cref.g2_neg(np.roll(g2, -1, axis=0)))
    dbl1 = cref.g1_double(g1)
    frp = fr_pow_oracle(fr, np.roll(fr, -1, axis=0))
    path = tmp_path / "vec.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", n))
        for a in (g1, g2, fr, gt, sg1, pw, add1, sub2, dbl1, frp):
            f.write(np.ascontiguousarray(a, dtype="<u8").tobytes())
    res = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "cpp api ok" in res.stdout


@pytest.mark.gpu
def test_cpp_api_multi_gpu_one_process(exe, tmp_path):
    """Library-level multi-GPU: ONE process, bn_b200_init_multi, the host-pointer
    bn_b200_pairing_batch shards the batch over every GPU of the box -- no torch, no NCCL on that path.  2^14 pairs per
    GPU (BASELINE config 5 on an 8-GPU box: 2^17).  Results must equal the single-GPU results everywhere and the oracle
    on a sample that includes both sides of every shard boundary."""
    import torch
    from oracle import cref
    gpus = torch.cuda.device_count()
    if gpus < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus N)")
    pairs = gpus << 14
    path = tmp_path / "sample.bin"
    res = subprocess.run([exe, "--multi", str(gpus), str(pairs), str(path)], capture_output=True, text=True)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(path, "rb").read()
    m = struct.unpack("<Q", raw[:8])[0]
    body = np.frombuffer(raw[8:], dtype="<u8")
    g1 = body[: m * 12].reshape(m, 12)
    g2 = body[m * 12: m * 36].reshape(m, 24)
    gt = body[m * 36:].reshape(m, 48)
    assert np.array_equal(gt, cref.pairing_batch(g1, g2, os.cpu_count() or 4))
    ratio = float(res.stdout.split("ratio ")[1].split()[0])
    assert ratio > 0.85 * gpus, res.stdout   # measured 0.96 x (2 GPUs) / 0.97 x (8 GPUs), profiles/; guards against serialisation
