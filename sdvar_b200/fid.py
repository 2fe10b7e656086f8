"""Checkpoint ingest + the FID sampling pipeline.

Reference procedure: load ``vae_ch160v4096z32.pth`` / ``var_d*.pth`` with ``load_state_dict(strict=True)``
(sdvar_colab_test.py:98-111, 170-172), sample 50 images per class with ``cfg=1.5, top_p=0.96, top_k=900, more_smooth=False``,
write them as PNG and pack the folder into one ``.npz`` (README.md:153, utils/misc.py:360-381: ``arr_0`` = (N,H,W,3) uint8).

Here the images never touch Python as floats: every generated batch is converted to uint8 HWC by one libsdvar pass
(``sdvar_image_to_u8``: trunc(x*255), what ``.mul_(255) ... astype(np.uint8)`` gives in the notebook), copied to pinned host
memory asynchronously while the next batch generates, and appended to the ``.npz`` array directly; PNG files are optional.
Under torch.distributed every rank samples its slice of the class list (``parallel.shard_range``) and writes its own shard."""
from __future__ import annotations

import os
from typing import Callable, Iterable, Optional, Sequence

import numpy as np
import torch

from . import _cabi, parallel


def load_checkpoints(vae=None, vae_ckpt: Optional[str] = None, **var_ckpts) -> None:
    """``load_checkpoints(vae, 'vae_ch160v4096z32.pth', draft=(draft_var, 'var_d16.pth'), target=(target_var, 'var_d30.pth'))``
    -- strict loads on the reference's checkpoint surface; the engines repack their bf16 copies lazily."""
    if vae is not None and vae_ckpt is not None:
        vae.load_state_dict(torch.load(vae_ckpt, map_location="cpu"), strict=True)
    for _, (model, path) in var_ckpts.items():
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=True)


@torch.no_grad()
def sample_fid_set(generate: Callable[[int, torch.Tensor, int], torch.Tensor], out_npz: str, classes: Iterable[int] = range(1000),
                   per_class: int = 50, batch: int = 100, seed: int = 0, device="cuda", png_dir: Optional[str] = None,
                   progress: Optional[Callable[[int, int], None]] = None) -> str:
    """Sample ``per_class`` images for every class and pack them DiT-style into ``out_npz`` (``arr_0``: (N,H,W,3) uint8).

    ``generate(B, label_B, g_seed) -> (B,3,H,W) fp32 in [0,1]`` is the model call, e.g.
    ``lambda B, lab, s: sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, g_seed=s, cfg=1.5, top_k=900, top_p=0.96)`` or
    ``lambda B, lab, s: var.autoregressive_infer_cfg(B, lab, g_seed=s, cfg=1.5, top_k=900, top_p=0.96)``.
    Returns the path written (``out_npz`` with a ``.rank{r}`` infix when running on several ranks)."""
    rank, world = parallel.world()
    labels = torch.tensor([c for c in classes for _ in range(per_class)], dtype=torch.int64)
    lo, hi = parallel.shard_range(labels.numel(), rank, world)
    labels = labels[lo:hi]
    N = labels.numel()
    arr = None
    host = [None, None]
    pending = None                      # (event, host buffer, start, count) of the batch still in flight
    if png_dir is not None:
        os.makedirs(png_dir, exist_ok=True)

    def drain(p):
        nonlocal arr
        ev, hb, s0, n = p
        ev.synchronize()
        a = hb[:n].numpy()
        if arr is None:
            arr = np.empty((N,) + a.shape[1:], dtype=np.uint8)
        arr[s0:s0 + n] = a
        if png_dir is not None:
            from PIL import Image
            for i in range(n):
                Image.fromarray(a[i]).save(os.path.join(png_dir, f"{lo + s0 + i:06d}.png"))

    for bi, s0 in enumerate(range(0, N, batch)):
        lab = labels[s0:s0 + batch].to(device)
        n = lab.numel()
        img = generate(n, lab, seed + (lo + s0))
        _, _, H, W = img.shape
        u8 = torch.empty(n, H, W, 3, dtype=torch.uint8, device=img.device)
        _cabi.image_to_u8(img.contiguous(), u8, hwc=True)
        slot = bi & 1
        if host[slot] is None:
            host[slot] = torch.empty(batch, H, W, 3, dtype=torch.uint8).pin_memory()
        if pending is not None:
            drain(pending)              # overlaps this batch's generation, which is already enqueued
        host[slot][:n].copy_(u8, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        pending = (ev, host[slot], s0, n)
        if progress is not None:
            progress(min(s0 + n, N), N)
    if pending is not None:
        drain(pending)
    path = out_npz if world == 1 else out_npz.replace(".npz", f".rank{rank}.npz")
    np.savez(path, arr_0=arr if arr is not None else np.empty((0, 0, 0, 3), np.uint8))
    return path


def merge_npz_shards(paths: Sequence[str], out_npz: str) -> str:
    """concatenate the per-rank shards in rank order (utils/misc.py:360-381 packs one folder into one file)"""
    np.savez(out_npz, arr_0=np.concatenate([np.load(p)["arr_0"] for p in paths], axis=0))
    return out_npz
