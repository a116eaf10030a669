"""Multi-GPU plumbing: candidate paths (or start/goal queries) are independent, so they shard by contiguous
ranges over the ranks with the map replicated per GPU; the only exchange is the final argmin of the best path
(path_generation/main.py:162-180) -- one 8-byte min-all-reduce over NCCL (gloo on CPU in the tests).

Key = (float32 bit pattern of the cost << 32) | global path index.  Costs are >= 0, so the bit pattern orders
like the value, the key is a non-negative int64, and ties resolve to the smaller index (the strict `<` of
main.py:175).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

KEY_EMPTY = np.iinfo(np.int64).max


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` units for `rank`; the first total % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f'bad rank/world {rank}/{world}')
    base, rem = divmod(int(total), world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def encode_key(cost: float, index: int) -> int:
    bits = int(np.float32(cost).view(np.uint32))
    return (bits << 32) | (int(index) & 0xFFFFFFFF)


def decode_key(key: int) -> Tuple[float, int]:
    key = int(key)
    return float(np.uint32((key >> 32) & 0xFFFFFFFF).view(np.float32)), key & 0xFFFFFFFF


def host_best_key(cost: np.ndarray, global_offset: int = 0) -> int:
    """Host-side packing of an already computed cost vector into the key (used for tiny batches and tests;
    the device path is Engine.best)."""
    cost = np.asarray(cost, dtype=np.float32)
    if cost.size == 0:
        return KEY_EMPTY
    i = int(np.argmin(cost))            # first minimum = smallest index on ties, like the key order
    if cost[i] > 0.0:                   # strictly positive and not NaN: the float order is the bit-pattern order
        return encode_key(cost[i], i + int(global_offset))
    keys = (cost.view(np.uint32).astype(np.uint64) << np.uint64(32)) | (
        (np.arange(cost.size, dtype=np.uint64) + np.uint64(global_offset)) & np.uint64(0xFFFFFFFF))
    return int(keys.min())


def global_best(local_key, group=None) -> Tuple[float, int]:
    """Min-all-reduce the per-rank key (1-element int64 tensor, on the GPU for NCCL) and decode it.
    Works without an initialised process group (single process) as the identity."""
    import torch
    import torch.distributed as dist
    key = local_key if torch.is_tensor(local_key) else torch.tensor([int(local_key)], dtype=torch.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MIN, group=group)
    return decode_key(int(key.item()))


def gather_costs(local_cost, group=None):
    """Optional second collective: all-gather of the per-path costs (4 bytes per path), equal shards."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_cost
    parts = [torch.empty_like(local_cost) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, local_cost, group=group)
    return torch.cat(parts)
