"""Data-parallel plumbing of the generation path: images are independent, so the batch is split
into contiguous per-rank slices, every rank runs its own draft+target+VQVAE replica with its own noise streams
(seed + rank), and the only collectives are ONE all_gather of the images and ONE all_reduce of the acceptance
counters per generation call.  Works with any torch.distributed backend (nccl on B200s, gloo in CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

STAT_KEYS = ("rounds", "target_passes", "draft_stages", "accepted_tokens", "rejected_tokens")


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """contiguous slice [lo, hi) of the global batch owned by `rank`; remainders go to the first ranks"""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_labels(labels: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None) -> torch.Tensor:
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_range(labels.shape[0], rank, world_size)
    return labels[lo:hi]


def rank_seed(g_seed: Optional[int], rank: Optional[int] = None) -> Optional[int]:
    """independent generator per rank so that any shard is reproducible by a single-GPU run of that slice"""
    if g_seed is None:
        return None
    return g_seed + (world()[0] if rank is None else rank)


def gather_images(img_local: torch.Tensor, global_batch: Optional[int] = None) -> torch.Tensor:
    """all_gather of (B_local,3,H,W) images into (B_global,3,H,W); ragged shards are padded to the largest one"""
    rank, w = world()
    if w == 1:
        return img_local
    sizes = [shard_range(global_batch, r, w)[1] - shard_range(global_batch, r, w)[0] for r in range(w)] if global_batch else None
    if sizes is None or len(set(sizes)) == 1:
        out = torch.empty((w * img_local.shape[0],) + tuple(img_local.shape[1:]), dtype=img_local.dtype, device=img_local.device)
        dist.all_gather_into_tensor(out, img_local.contiguous())
        return out
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(img_local.shape[1:]), dtype=img_local.dtype, device=img_local.device)
    pad[:img_local.shape[0]] = img_local
    bufs = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], 0)


def reduce_stats(stats: Dict[str, int], device) -> Dict[str, int]:
    """sum of the acceptance counters over ranks (<= 64 bytes on the wire)"""
    rank, w = world()
    if w == 1:
        return {k: int(stats[k]) for k in STAT_KEYS}
    t = torch.tensor([int(stats[k]) for k in STAT_KEYS], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return {k: int(v) for k, v in zip(STAT_KEYS, t.tolist())}
