"""Multi-GPU plumbing: candidate paths (or start/goal queries) are independent, so they shard by contiguous
ranges over the ranks with the map replicated per GPU; the only exchange is the final argmin of the best path
(path_generation/main.py:162-180) -- one 8-byte min-all-reduce over NCCL (gloo on CPU in the tests).

Key = (img(float32 cost) << 31) | global path index (31 bits), img = order-preserving image of the float32 bit
pattern: all bits flipped for a negative value, top bit set otherwise, NaN -> 0xFFFFFFFF.  The integer order of the
keys is the float order of the costs for every value (negative costs are reachable through negative layer weights),
a NaN never wins, ties resolve to the smaller index (the strict `<` of main.py:175), and every key is a non-negative
int64 <= KEY_EMPTY, so the device's unsigned atomicMin and the signed MIN of the all-reduce agree.
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


def _image(cost) -> np.ndarray:
    """order-preserving uint32 image of float32 values (uam_best_key in csrc/uam_internal.cuh)"""
    c = np.atleast_1d(np.asarray(cost, dtype=np.float32))
    bits = c.view(np.uint32)
    img = np.where(bits & np.uint32(0x80000000), ~bits, bits | np.uint32(0x80000000))
    return np.where(np.isnan(c), np.uint32(0xFFFFFFFF), img).astype(np.uint32)


def encode_key(cost: float, index: int) -> int:
    if not 0 <= int(index) < 2 ** 31:
        raise ValueError('global path index must fit 31 bits')
    return (int(_image(cost)[0]) << 31) | int(index)


def decode_key(key: int) -> Tuple[float, int]:
    """(cost, global index); KEY_EMPTY (no candidate) decodes to (nan, 2^31 - 1)"""
    key = int(key)
    img = (key >> 31) & 0xFFFFFFFF
    bits = (img ^ 0x80000000) if img & 0x80000000 else (~img & 0xFFFFFFFF)
    return float(np.uint32(bits).view(np.float32)), key & 0x7FFFFFFF


def host_best_key(cost: np.ndarray, global_offset: int = 0) -> int:
    """Host-side packing of an already computed cost vector into the key (used for tiny batches, the host-buffer
    entry points and tests; the device path is Engine.best)."""
    cost = np.asarray(cost, dtype=np.float32).ravel()
    if cost.size == 0:
        return KEY_EMPTY
    if int(global_offset) < 0 or int(global_offset) + cost.size > 2 ** 31 - 1:
        raise ValueError('global path index must fit 31 bits')
    if not np.isnan(cost).any():
        i = int(np.argmin(cost))        # first minimum = smallest index on ties, like the key order (-0.0 < +0.0 aside)
        if cost[i] != 0.0:
            return encode_key(cost[i], i + int(global_offset))
    keys = (_image(cost).astype(np.uint64) << np.uint64(31)) | (np.arange(cost.size, dtype=np.uint64) + np.uint64(global_offset))
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


def attach_peer_group(engine, group=None) -> int:
    """Give `engine` the symmetric blocks of every rank of the box: all-gather the 64-byte CUDA IPC handles
    (Engine.peer_handle) over torch.distributed (any backend; the handles are plain bytes) and map them
    (Engine.attach_peers).  After this, RasterMap.score_paths_best / Engine.best_allreduce return the min over all ranks
    without any NCCL call on the data path.  Returns the world size (1 = nothing to attach)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    handles = [None] * world
    dist.all_gather_object(handles, engine.peer_handle(), group=group)
    engine.attach_peers(rank, handles)
    dist.barrier(group=group)           # every rank has mapped every block before the first exchange
    return world
