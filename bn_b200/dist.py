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


class FusedGather:
    """Peer-mapped gather buffer for `bn_b200_pairing_batch_gather_dev`: every rank allocates world * n Gt slots in
    symmetric (peer-accessible) device memory; the final-exponentiation kernel of rank r stores its results straight into
    slot range [r * n, (r + 1) * n) of EVERY rank's buffer over NVLink, so no separate all-gather runs.

    Uses torch.distributed._symmetric_memory for the allocation / handle exchange and its barrier for the cross-rank
    ordering; raises if peer access is not available (the caller then keeps the NCCL all-gather)."""

    def __init__(self, n_per_rank: int, device, group=None):
        import ctypes

        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("FusedGather supports up to 8 peers (one NVSwitch domain)")
        self.n = n_per_rank
        self.buf = symm_mem.empty((self.world * n_per_rank, 48), dtype=torch.int64, device=device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != self.world or not all(ptrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer pointers")
        self._ptrs = (ctypes.c_void_p * self.world)(*ptrs)

    def pairing_batch(self, lib, d_g1, d_g2, stream_ptr):
        """Enqueue lines + Miller + final exponentiation with the fused peer-store epilogue for this rank's n pairs."""
        import ctypes
        from . import _lib
        _lib.check(lib.bn_b200_pairing_batch_gather_dev(ctypes.c_void_p(d_g1.data_ptr()), ctypes.c_void_p(d_g2.data_ptr()),
                                                        self._ptrs, self.world, self.rank, ctypes.c_size_t(self.n), stream_ptr))

    def barrier(self):
        """All ranks' kernels up to here have completed and their peer stores are visible."""
        self.handle.barrier()

    def result(self):
        return self.buf
