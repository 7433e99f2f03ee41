"""Instance sharding across the GPUs of one node.

MPC instances are independent, so the data path has NO collective: rank r solves the contiguous slice
`shard_range(B, r, world)` of the batch on its own GPU.  The single exchange of the path is the final
gather of the packed result buffers (NCCL all-gather over NVLink on the GPU box; the same code runs on
gloo/CPU tensors in tests/test_sharding_gloo.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import _capi


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous slice; the first B % world ranks take one extra instance."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(int(B), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def packed_layout(B: int, H: int, fields=None) -> Tuple[Dict[str, Tuple[int, int]], int]:
    """name -> (byte offset, byte size) of each field's 256-byte aligned slab, and the total size.
    Same layout as BatchedMPC.alloc_device_outputs."""
    spec = _capi.output_spec(H)
    fields = list(spec) if fields is None else list(fields)
    offs, total = {}, 0
    for name in fields:
        shp, dt = spec[name]
        nbytes = B * int(np.prod(shp, dtype=np.int64)) * np.dtype(dt).itemsize
        offs[name] = (total, nbytes)
        total = (total + nbytes + 255) // 256 * 256
    return offs, max(total, 256)


def gather_packed(packed, B_total: int, H: int, fields=None, group=None):
    """ONE collective: all-gather every rank's packed result buffer, then return typed views of the
    whole batch in instance order: dict name -> list of per-rank tensors (views, no copy).
    Ranks may hold ragged shards (shard_range); buffers are padded to the largest shard's size."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(B_total, r, world) for r in range(world)]
    cap = max(packed_layout(hi - lo, H, fields)[1] for lo, hi in sizes)
    mine = packed
    if packed.numel() != cap:
        mine = torch.zeros(cap, dtype=torch.uint8, device=packed.device)
        mine[: packed.numel()] = packed
    out = torch.empty(world * cap, dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    spec = _capi.output_spec(H)
    names = list(spec) if fields is None else list(fields)
    views: Dict[str, List] = {k: [] for k in names}
    for r, (lo, hi) in enumerate(sizes):
        offs, _ = packed_layout(hi - lo, H, names)
        chunk = out[r * cap:(r + 1) * cap]
        for k in names:
            o, nb = offs[k]
            shp, dt = spec[k]
            views[k].append(chunk[o:o + nb].view(getattr(torch, dt)).view((hi - lo,) + shp))
    return views


def concat_views(views) -> Dict[str, "np.ndarray"]:
    """Host copy of gather_packed's views as whole-batch numpy arrays (instance order)."""
    import torch

    return {k: torch.cat(v, dim=0).cpu().numpy() for k, v in views.items()}
