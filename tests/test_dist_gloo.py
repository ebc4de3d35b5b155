"""World-size-2 gloo test (CPU) of the multi-GPU path's host logic: contiguous sharding + all-gather of Gt.
The per-rank compute is injected (the CPU oracle stands in for the GPU engine, which has no CPU fallback);
the -m gpu suite and bench.py --gpus N run the same code over NCCL with the real kernels."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bn_b200 import dist as bdist
    from oracle import cref
    from tests import util
    g1, g2 = util.synth_pairs(0xD157, n, 2)
    calls = []

    def compute(a, b):
        calls.append(len(a))
        return cref.pairing_batch(a, b, 2)

    out = bdist.pairing_batch_sharded(g1, g2, compute=compute)
    lo, hi = bdist.shard_range(n, rank, world)
    q.put((rank, out.tobytes(), calls, (lo, hi)))
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    from bn_b200.dist import shard_range
    for n in (0, 1, 7, 16, 17, 1 << 14):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_sharded_pairing_world2_gloo():
    from oracle import cref
    from tests import util
    n, world = 13, 2  # ragged: shards of 7 and 6
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    g1, g2 = util.synth_pairs(0xD157, n, 2)
    want = cref.pairing_batch(g1, g2, 2).tobytes()
    for rank, got, calls, span in res:
        assert got == want, rank               # sharded == single-process == oracle
        assert calls == [span[1] - span[0]]    # each rank computed only its own shard
