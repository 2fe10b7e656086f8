"""Model builders with the reference's signatures (reference models/__init__.py:16-97)."""
from typing import Tuple

import torch
import torch.nn as nn

from .quant import VectorQuantizer2
from .var import SDVAR, VAR
from .vqvae import VQVAE

__all__ = ["VQVAE", "VAR", "SDVAR", "VectorQuantizer2", "build_vae_var", "build_vae_var_speculative_decoding"]


def _make_var(vae, device, patch_nums, num_classes, depth, shared_aln, attn_l2_norm, flash_if_available, fused_if_available,
              init_adaln, init_adaln_gamma, init_head, init_std) -> VAR:
    var = VAR(vae_local=vae, num_classes=num_classes, depth=depth, embed_dim=64 * depth, num_heads=depth, drop_rate=0.0,
              attn_drop_rate=0.0, drop_path_rate=0.1 * depth / 24, norm_eps=1e-6, shared_aln=shared_aln, cond_drop_rate=0.1,
              attn_l2_norm=attn_l2_norm, patch_nums=patch_nums, flash_if_available=flash_if_available,
              fused_if_available=fused_if_available).to(device)
    var.init_weights(init_adaln=init_adaln, init_adaln_gamma=init_adaln_gamma, init_head=init_head, init_std=init_std)
    return var


def build_vae_var(device, patch_nums=(1, 2, 3, 4, 5, 6, 8, 10, 13, 16), V=4096, Cvae=32, ch=160, share_quant_resi=4,
                  num_classes=1000, depth=16, shared_aln=False, attn_l2_norm=True, flash_if_available=True,
                  fused_if_available=True, init_adaln=0.5, init_adaln_gamma=1e-5, init_head=0.02, init_std=-1) -> Tuple[VQVAE, VAR]:
    """width = 64*depth, heads = depth (reference models/__init__.py:26-27).  Unlike the reference this does NOT
    monkey-patch ``reset_parameters`` process-wide (models/__init__.py:31-32): the VQVAE keeps torch's default init
    until a checkpoint or ``sdvar_b200.weights.vqvae_state_dict`` is loaded, instead of uninitialised memory."""
    vae = VQVAE(vocab_size=V, z_channels=Cvae, ch=ch, test_mode=True, share_quant_resi=share_quant_resi, v_patch_nums=patch_nums).to(device)
    var = _make_var(vae, device, patch_nums, num_classes, depth, shared_aln, attn_l2_norm, flash_if_available, fused_if_available,
                    init_adaln, init_adaln_gamma, init_head, init_std)
    return vae, var


def build_vae_var_speculative_decoding(device, patch_nums=(1, 2, 3, 4, 5, 6, 8, 10, 13, 16), V=4096, Cvae=32, ch=160,
                                       share_quant_resi=4, num_classes=1000, depth_draft=16, depth_target=30, shared_aln=False,
                                       attn_l2_norm=True, flash_if_available=True, fused_if_available=True, init_adaln=0.5,
                                       init_adaln_gamma=1e-5, init_head=0.02, init_std=-1, similarity_thresh=0.8,
                                       shared_aln_target=None):
    """-> (vae, draft, target, sdvar)  (reference models/__init__.py:51-97).  ``shared_aln_target`` (extension) lets
    the target use shared adaLN while the draft does not (d36 at 512 px, README.md:142-144; fixes D10)."""
    vae = VQVAE(vocab_size=V, z_channels=Cvae, ch=ch, test_mode=True, share_quant_resi=share_quant_resi, v_patch_nums=patch_nums).to(device)
    draft = _make_var(vae, device, patch_nums, num_classes, depth_draft, shared_aln, attn_l2_norm, flash_if_available,
                      fused_if_available, init_adaln, init_adaln_gamma, init_head, init_std)
    target = _make_var(vae, device, patch_nums, num_classes, depth_target,
                       shared_aln if shared_aln_target is None else shared_aln_target, attn_l2_norm, flash_if_available,
                       fused_if_available, init_adaln, init_adaln_gamma, init_head, init_std)
    return vae, draft, target, SDVAR(draft, target, similarity_thresh)
