"""Concurrency and multi-GPU behaviour of the C ABI on hardware: calls on different streams (shared scratch),
calls from several host threads, re-binding to another device, library-level multi-GPU from one process, and the
fused peer-store gather across ranks checked against the ORACLE (VERDICT r1 missing #6)."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from oracle import cref
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bn():
    import bn_b200
    bn_b200.init(0)
    return bn_b200


def test_two_streams_share_the_scratch_safely(bn):
    """Two *_dev pairing calls enqueued back to back on DIFFERENT streams (and a third on the library's own stream, NULL):
    the line / flag scratch is ordered between them by the library; every result must be right."""
    import torch
    lib = bn.load()
    dev = torch.device("cuda", 0)
    sets = []
    for i, n in enumerate((4000, 3000, 1500)):
        g1, g2 = util.synth_pairs(0x57E0 + i, 40)
        g1, g2 = np.tile(g1, (n // 40 + 1, 1))[:n], np.tile(g2, (n // 40 + 1, 1))[:n]
        sets.append((g1, g2, torch.from_numpy(g1.view(np.int64)).to(dev), torch.from_numpy(g2.view(np.int64)).to(dev),
                     torch.zeros((n, 48), dtype=torch.int64, device=dev)))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev), None]
    for rep in range(3):
        for (g1, g2, d1, d2, out), st in zip(sets, streams):
            sp = ctypes.c_void_p(st.cuda_stream) if st is not None else None
            assert lib.bn_b200_pairing_batch_dev(ctypes.c_void_p(d1.data_ptr()), ctypes.c_void_p(d2.data_ptr()),
                                                 ctypes.c_void_p(out.data_ptr()), ctypes.c_size_t(len(g1)), sp) == 0
    torch.cuda.synchronize()
    assert lib.bn_b200_device_error(0) == 0   # also drains the library's own stream
    for g1, g2, d1, d2, out in sets:
        want40 = cref.pairing_batch(g1[:40], g2[:40], 8)
        got = out.cpu().numpy().view(np.uint64)
        assert np.array_equal(got, np.tile(want40, (len(g1) // 40 + 1, 1))[:len(g1)])


def test_host_calls_from_many_threads(bn):
    """Host-pointer calls from 4 host threads at once (none of which ever called cudaSetDevice): per-device lock,
    double-buffered staging, the device guard.  All results bit-exact."""
    g1, g2 = util.synth_pairs(0x7EAD, 48)
    want = cref.pairing_batch(g1, g2, 8)
    k = util.synth_scalars(0x7EAE, 48)
    want_mul = cref.g1_mul_batch(g1, k, 8)
    errs = []

    def work(t):
        try:
            for r in range(6):
                lo = (7 * t + 5 * r) % 20
                assert np.array_equal(bn.pairing_batch(g1[lo:], g2[lo:]), want[lo:])
                assert np.array_equal(bn.g1_mul_batch(g1[lo:], k[lo:]), want_mul[lo:])
        except BaseException as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_rebind_and_multi_init(bn):
    """bn_b200_init on another device releases the old binding; bn_b200_init_multi binds every GPU and the host-pointer
    pairing call is sharded over them (one process).  With one GPU this still exercises teardown / re-creation."""
    import torch
    lib = bn.load()
    gpus = torch.cuda.device_count()
    n = 2 * 2048 * max(gpus, 1) + 37     # above the minimum shard size for every device, ragged tail
    g1, g2 = util.synth_pairs(0x3B1D, 64)
    g1, g2 = np.tile(g1, (n // 64 + 1, 1))[:n].copy(), np.tile(g2, (n // 64 + 1, 1))[:n].copy()
    want64 = cref.pairing_batch(g1[:64], g2[:64], 8)
    want = np.tile(want64, (n // 64 + 1, 1))[:n]
    try:
        assert lib.bn_b200_shutdown() == 0 and lib.bn_b200_device_count() == 0
        out = np.zeros((4, 48), dtype=np.uint64)
        assert lib.bn_b200_pairing_batch(g1.ctypes.data_as(ctypes.c_void_p), g2.ctypes.data_as(ctypes.c_void_p),
                                         out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(4)) == -1   # not bound: ENODEV
        bn.init_multi(0)
        assert lib.bn_b200_device_count() == gpus
        assert np.array_equal(bn.pairing_batch(g1, g2), want)                   # pageable buffers: staged, sharded
        h1 = torch.from_numpy(g1.view(np.int64)).pin_memory()
        h2 = torch.from_numpy(g2.view(np.int64)).pin_memory()
        ho = torch.zeros((n, 48), dtype=torch.int64).pin_memory()
        assert lib.bn_b200_pairing_batch(ctypes.c_void_p(h1.data_ptr()), ctypes.c_void_p(h2.data_ptr()),
                                         ctypes.c_void_p(ho.data_ptr()), ctypes.c_size_t(n)) == 0  # pinned: zero-copy epilogue per device
        assert np.array_equal(ho.numpy().view(np.uint64), want)
        k = util.synth_scalars(0x3B1E, 64)
        kk = np.tile(k, (n // 64 + 1, 1))[:n].copy()
        assert np.array_equal(bn.pairing_pow_batch(g1, g2, kk)[:64], cref.gt_pow_batch(want64, k, 8))
        assert np.array_equal(bn.g2_mul_batch(g2, kk)[:64], cref.g2_mul_batch(g2[:64], k, 8))
        if gpus > 1:
            bn.init(gpus - 1)                                                    # re-bind to a single, different device
            assert lib.bn_b200_device_count() == 1
            assert np.array_equal(bn.pairing_batch(g1[:100], g2[:100]), want[:100])
    finally:
        bn.init(0)
    assert np.array_equal(bn.pairing_batch(g1[:50], g2[:50]), want[:50])


RANK_SCRIPT = r'''
import os, sys, ctypes
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["BN_ROOT"])
import bn_b200
from bn_b200.dist import FusedGather
from oracle import cref
from tests import util
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
lib = bn_b200.init(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = 1000 + 37
g1, g2 = util.synth_pairs(0xFA5E + rank, 64)           # every rank its own inputs
g1, g2 = np.tile(g1, (n // 64 + 1, 1))[:n], np.tile(g2, (n // 64 + 1, 1))[:n]
dev = torch.device("cuda", rank)
d1, d2 = torch.from_numpy(g1.view(np.int64)).to(dev), torch.from_numpy(g2.view(np.int64)).to(dev)
fg = FusedGather(n, dev)
st = torch.cuda.Stream(device=dev)
with torch.cuda.stream(st):
    fg.pairing_batch(lib, d1, d2, ctypes.c_void_p(st.cuda_stream))
    fg.barrier()
torch.cuda.synchronize()
got = fg.result().cpu().numpy().view(np.uint64)
# the oracle's answer for EVERY rank's slot, recomputed locally from the seeds
ok = True
for r in range(world):
    a, b = util.synth_pairs(0xFA5E + r, 64)
    want = np.tile(cref.pairing_batch(a, b, 4), (n // 64 + 1, 1))[:n]
    ok = ok and np.array_equal(got[r * n:(r + 1) * n], want)
# the NCCL / numpy sharded helper against the oracle as well
full1 = np.concatenate([np.tile(util.synth_pairs(0xFA5E + r, 64)[0], (2, 1)) for r in range(world)])
full2 = np.concatenate([np.tile(util.synth_pairs(0xFA5E + r, 64)[1], (2, 1)) for r in range(world)])
from bn_b200.dist import pairing_batch_sharded
sh = pairing_batch_sharded(full1, full2, device=rank)
ok = ok and np.array_equal(sh, cref.pairing_batch(full1, full2, 4))
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.barrier()
dist.destroy_process_group()
print("RANK %d %s" % (rank, "OK" if ok else "MISMATCH"), flush=True)
sys.exit(0 if int(flag.item()) == 0 else 1)
'''


def test_fused_gather_two_ranks_vs_oracle(bn, tmp_path):
    """FusedGather at world > 1: every slot of every rank's gathered buffer equals the oracle (not merely the NCCL path)."""
    import torch
    gpus = torch.cuda.device_count()
    if gpus < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = min(gpus, 4)
    script = tmp_path / "rank.py"
    script.write_text(RANK_SCRIPT)
    env = dict(os.environ, BN_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("OK") == world
