"""Replica data-parallelism over independent prompts: one process per GPU, image i -> rank i % W,
every rank holds a full UNet replica, ONE broadcast of the weights at init (so replicas are bit-identical), no
per-step collective, optional final gather of the latents. torch.distributed (NCCL on GPUs, gloo in CPU tests) is
plumbing; the reference itself has no distributed code (it loops over prompts: examples/text_to_mscoco.py:54-62)."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin partition: item i belongs to rank i % world."""
    return list(range(rank, n_items, world))


def broadcast_state_dict(sd: Optional[Dict[str, torch.Tensor]], specs, device, src: int = 0,
                         dtype=torch.float16, bucket_elems: int = 1 << 28) -> Dict[str, torch.Tensor]:
    """Rank `src` passes its state dict; everybody returns an identical copy. Tensors travel in large flat buckets
    (few big NCCL broadcasts over NVLink rather than ~1700 small ones). `specs` = [(key, shape, kind)] fixes the
    order and shapes on ranks that have no weights yet."""
    rank = dist.get_rank()
    out: Dict[str, torch.Tensor] = {}
    bucket: List = []
    n = 0

    def flush():
        nonlocal bucket, n
        if not bucket:
            return
        flat = torch.empty(n, dtype=dtype, device=device)
        if rank == src:
            off = 0
            for key, shape in bucket:
                k = sd[key].numel()
                flat[off:off + k].copy_(sd[key].reshape(-1).to(device=device, dtype=dtype))
                off += k
        dist.broadcast(flat, src=src)
        off = 0
        for key, shape in bucket:
            k = 1
            for d in shape:
                k *= d
            out[key] = flat[off:off + k].view(shape)
            off += k
        bucket, n = [], 0

    for key, shape, _ in specs:
        k = 1
        for d in shape:
            k *= d
        if n + k > bucket_elems:
            flush()
        bucket.append((key, tuple(shape)))
        n += k
    flush()
    return out


def gather_latents(local: torch.Tensor, n_items: int, dst: int = 0) -> Optional[torch.Tensor]:
    """All ranks contribute their (n_local,4,h,w) latents; rank `dst` gets them re-ordered to item order."""
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_items + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_items,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = bufs[r][: len(idx)]
    return out
