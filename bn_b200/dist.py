"""Multi-GPU sharding of a pairing batch: one process per GPU, contiguous shards, one all-gather of Gt.

SURVEY.md section 8(e): pairings are independent, so the batch is split into contiguous ranges of ceil(n/G)
pairs, each rank runs the two kernels on its shard, and the only exchange is gathering the 384-byte results
(NCCL over NVLink on GPUs; the same code path runs over gloo in the CPU tests).  No reduction, no all-to-all.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank `rank`; shards differ in size by at most one block of ceil(n/world)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def pairing_batch_sharded(g1: np.ndarray, g2: np.ndarray, compute=None, device=None) -> np.ndarray:
    """e(g1[i], g2[i]) for the whole batch, computed cooperatively by all ranks of the default process group;
    every rank returns the full [n, 48] result.

    `compute(g1_shard, g2_shard) -> gt_shard` defaults to the GPU engine bound to this rank's device
    (bn_b200.pairing_batch); tests inject a CPU stand-in to exercise the sharding/gather logic under gloo.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = len(g1)
    if compute is None:
        import bn_b200
        dev_index = device if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
        compute = lambda a, b: bn_b200.pairing_batch(a, b, device=dev_index)  # noqa: E731
    lo, hi = shard_range(n, rank, world)
    local = compute(g1[lo:hi], g2[lo:hi]) if hi > lo else np.zeros((0, 48), dtype=np.uint64)
    if world == 1:
        return local
    per = (n + world - 1) // world
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((per, 48), dtype=torch.int64, device=dev)  # padded so every rank contributes the same size
    if hi > lo:
        buf[: hi - lo] = torch.from_numpy(np.ascontiguousarray(local).view(np.int64)).to(dev)
    out = torch.empty((world * per, 48), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, buf)
    return out.cpu().numpy().view(np.uint64)[:n]
