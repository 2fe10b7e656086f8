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


class GatherBuffer:
    """The image all_gather of a generation call, done the cheap way: every rank converts its fp32 images to
    uint8 with one libsdvar pass (192 KiB instead of 768 KiB per 256 px image) straight into its slot of ONE preallocated
    (world*B,3,H,W) uint8 buffer, and the collective runs in place on that buffer (``all_gather_into_tensor`` with the input
    aliasing the rank's slot).  No host synchronisation; reused across calls.  CPU tensors (gloo tests) take a torch path."""

    def __init__(self, world_size: int, B_local: int, chw, device):
        self.w, self.B = world_size, B_local
        self.out = torch.empty((world_size * B_local,) + tuple(chw), dtype=torch.uint8, device=device)

    def gather(self, img_local: torch.Tensor) -> torch.Tensor:
        rank, w = world()
        mine = self.out[rank * self.B:(rank + 1) * self.B]
        if img_local.is_cuda:
            from . import _cabi
            _cabi.image_to_u8(img_local.contiguous(), mine)
        else:
            mine.copy_((img_local.clamp(0, 1) * 255).to(torch.uint8))
        if w > 1:
            dist.all_gather_into_tensor(self.out, mine)
        return self.out


def reduce_stats_async(stats: Dict[str, int], device) -> torch.Tensor:
    """sum of the acceptance counters over ranks, left ON THE DEVICE (no host sync inside the step); read it with
    ``stats_from_tensor`` after the timed region"""
    t = torch.tensor([int(stats[k]) for k in STAT_KEYS], dtype=torch.int64).to(device, non_blocking=True)
    if world()[1] > 1:
        dist.all_reduce(t)
    return t


def stats_from_tensor(t: torch.Tensor) -> Dict[str, int]:
    return {k: int(v) for k, v in zip(STAT_KEYS, t.tolist())}


def reduce_stats(stats: Dict[str, int], device) -> Dict[str, int]:
    """sum of the acceptance counters over ranks (<= 64 bytes on the wire)"""
    rank, w = world()
    if w == 1:
        return {k: int(stats[k]) for k in STAT_KEYS}
    t = torch.tensor([int(stats[k]) for k in STAT_KEYS], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return {k: int(v) for k, v in zip(STAT_KEYS, t.tolist())}
