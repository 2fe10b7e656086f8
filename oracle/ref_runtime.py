"""Run the REAL reference (unmodified modules from oracle/_ref/reference_path.tar.gz, see oracle/build_ref.py) on the host
CPU.  BASELINE INFRASTRUCTURE: only bench.py's reference arm / cpu_baseline leg and tests may import this.

The reference picks 'cuda' whenever a GPU is visible (dist.py:12) and hard-codes ``torch.device('cuda:0')`` inside sd_test3
(models/var.py:737, defect D11); on a box with a GPU that would split tensors across devices, so its device selection is set
to 'cpu' (the module global dist.py:64 returns) and ``torch.device`` is shimmed to the CPU *inside the reference's var module
only* while its functions run.  Nothing else is patched: the loops, blocks, sampler and VQVAE are the reference's own code."""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tarfile
import tempfile
import time
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "reference_path.tar.gz")
_MODELS = None


def available() -> bool:
    return os.path.exists(ARCHIVE)


def load():
    """import the reference's ``models`` package from the archive (once per process)"""
    global _MODELS
    if _MODELS is None:
        if not available():
            raise RuntimeError(f"{ARCHIVE} is missing: run `python oracle/build_ref.py` where /root/reference exists")
        root = tempfile.mkdtemp(prefix="sdvar_ref_")
        with tarfile.open(ARCHIVE) as tar:
            tar.extractall(root)
        sys.path.insert(0, root)
        with contextlib.redirect_stdout(io.StringIO()):
            import dist as RD                       # the reference's device selection (dist.py:12): pick the CPU for this arm
            setattr(RD, "__device", "cpu")
            import models as R                      # noqa: the reference prints banners on import
        _MODELS = R
    return _MODELS


@contextlib.contextmanager
def _cpu_device_shim():
    R = load()
    import models.var as RV
    real = RV.torch.device
    RV.torch.device = lambda *a, **k: real("cpu")
    try:
        with contextlib.redirect_stdout(io.StringIO()):     # the reference prints per-stage debug lines
            yield R
    finally:
        RV.torch.device = real


def build_models(patch_nums, depth_draft: int, depth_target: int, sd_draft, sd_target, sd_vae, shared_aln_target: bool = False):
    """reference VQVAE + draft VAR + target VAR + SDVAR on the CPU, loaded (strict) with the given state dicts"""
    R = load()
    with contextlib.redirect_stdout(io.StringIO()):
        vae = R.VQVAE(vocab_size=4096, z_channels=32, ch=160, test_mode=True, share_quant_resi=4, v_patch_nums=patch_nums)
        mk = lambda depth, saln: R.VAR(vae_local=vae, depth=depth, embed_dim=64 * depth, num_heads=depth, attn_l2_norm=True,
                                       patch_nums=patch_nums, shared_aln=saln, flash_if_available=False, fused_if_available=False).eval()
        draft, target = mk(depth_draft, False), mk(depth_target, shared_aln_target)
    r = vae.load_state_dict(sd_vae, strict=False)
    assert not r.unexpected_keys and all(k.startswith(("encoder.", "quant_conv.")) for k in r.missing_keys)
    draft.load_state_dict(sd_draft, strict=True)
    target.load_state_dict(sd_target, strict=True)
    with _cpu_device_shim():
        sd = R.SDVAR(draft, target)
    return vae, draft, target, sd


def time_entry(fn, warmup: int, steps: int):
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        with _cpu_device_shim(), torch.no_grad():
            fn(i)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts)
