"""Shared test helpers: conversions between the big-int oracle's tuples and #[repr(C)] byte images,
and deterministic synthetic inputs.  Test infrastructure only."""
import json
import os

import numpy as np

from oracle import bn_oracle as o
from oracle import cref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def load_vectors(name, limit=None):
    with open(os.path.join(GOLDEN, name)) as f:
        lines = [l.strip() for l in f if l.strip() and not l.startswith("#")]
    return lines[:limit] if limit else lines


def words(b: bytes) -> np.ndarray:
    return np.frombuffer(b, dtype="<u8").astype(np.uint64)


def fr_img(k: int) -> np.ndarray:
    return words(o.fr_to_bytes(k))


def g1_img(p) -> np.ndarray:
    return words(o.g1_to_bytes(p))


def g2_img(p) -> np.ndarray:
    return words(o.g2_to_bytes(p))


def gt_img(f) -> np.ndarray:
    return words(o.gt_to_bytes(f))


def img_g1(w):
    return o.g1_from_bytes(np.asarray(w, dtype="<u8").tobytes())


def img_g2(w):
    return o.g2_from_bytes(np.asarray(w, dtype="<u8").tobytes())


def img_gt(w):
    return o.gt_from_bytes(np.asarray(w, dtype="<u8").tobytes())


def synth_scalars(seed: int, n: int, start: int = 0) -> np.ndarray:
    """[n,4] Montgomery Fr images of synth_scalar(seed, i)."""
    return np.stack([fr_img(o.synth_scalar(seed, start + i)) for i in range(n)])


def synth_pairs(seed: int, n: int, threads: int = 8):
    """n random (G1,G2) Jacobian pairs, z != 1: P_i = G1::one()*a_i, Q_i = G2::one()*b_i
    (mirrors G::random, reference src/groups/mod.rs:220-222).  Built with the C oracle."""
    a = synth_scalars(seed, n, 0)
    b = synth_scalars(seed ^ 0x5555, n, 1 << 20)
    g1 = cref.g1_mul_batch(np.repeat(cref.g1_generator(), n, axis=0), a, threads)
    g2 = cref.g2_mul_batch(np.repeat(cref.g2_generator(), n, axis=0), b, threads)
    return g1, g2


def edge_case_pairs():
    """(G1,G2) images exercising the reference's special cases."""
    gen1, gen2 = cref.g1_generator()[0], cref.g2_generator()[0]
    inf1, inf2 = g1_img(o.g_zero(o.FQ)), g2_img(o.g_zero(o.FQ2))
    p5 = cref.g1_mul_batch(gen1[None], fr_img(5)[None])[0]
    q7 = cref.g2_mul_batch(gen2[None], fr_img(7)[None])[0]
    # non-canonical infinity: (x, y, 0) with arbitrary x, y
    weird_inf1 = p5.copy(); weird_inf1[8:12] = 0
    weird_inf2 = q7.copy(); weird_inf2[16:24] = 0
    g1s = [gen1, inf1, gen1, p5, weird_inf1, p5, cref.g1_normalize(p5[None])[0], cref.g1_neg(p5[None])[0]]
    g2s = [gen2, gen2, inf2, q7, q7, weird_inf2, cref.g2_normalize(q7[None])[0], q7]
    return np.stack(g1s), np.stack(g2s)
