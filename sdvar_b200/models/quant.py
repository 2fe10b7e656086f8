"""VectorQuantizer2, inference side -- same interface and state-dict keys as the reference's
models/quant.py (embedding, quant_resi.qresi_ls.{i}, ema_vocab_hit_SV), with the per-stage f_hat update
done by the fused K5 kernel (``sdvar_vq_next_input``).

In scope: ``embedding``, ``get_next_autoregressive_input``, ``embed_to_fhat`` /
``idxBl_to_var_input`` built on the same kernel, and the encode side ``f_to_idxBl_or_fhat`` (SURVEY.md 8f #3:
nearest-code kernel + the same K5 step keeping the reference's running residual).  The VAE-training ``forward``
is out of scope and raises.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import _cabi


class _PhiParams(nn.Module):
    """Parameter holder of one Phi conv (reference: class Phi(nn.Conv2d), models/quant.py:199-206):
    out = (1-r)*h + r*conv3x3(h), r = 0.5."""

    def __init__(self, Cvae: int, resi_ratio: float):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(Cvae, Cvae, 3, 3))
        self.bias = nn.Parameter(torch.empty(Cvae))
        self.resi_ratio = abs(resi_ratio)


class _PhiBank(nn.Module):
    """Shared / partially-shared / non-shared Phi selection (models/quant.py:209-243).  ``index(at)`` returns
    which Phi serves relative depth ``at = si/(SN-1)``: argmin |ticks - at|."""

    def __init__(self, n: int, Cvae: int, resi_ratio: float, mode: str):
        super().__init__()
        self.mode = mode
        if mode == "shared":
            self.qresi = _PhiParams(Cvae, resi_ratio)
        else:
            self.qresi_ls = nn.ModuleList([_PhiParams(Cvae, resi_ratio) for _ in range(n)])
        K = n
        self.ticks = np.linspace(1 / 3 / K, 1 - 1 / 3 / K, K) if K == 4 else np.linspace(1 / 2 / K, 1 - 1 / 2 / K, K)

    def index(self, at_from_0_to_1: float) -> int:
        return 0 if self.mode == "shared" else int(np.argmin(np.abs(self.ticks - at_from_0_to_1)).item())

    def __getitem__(self, at_from_0_to_1: float) -> _PhiParams:
        return self.qresi if self.mode == "shared" else self.qresi_ls[self.index(at_from_0_to_1)]


class VectorQuantizer2(nn.Module):
    def __init__(self, vocab_size, Cvae, using_znorm=False, beta: float = 0.25, default_qresi_counts=0, v_patch_nums=None,
                 quant_resi=0.5, share_quant_resi=4):
        super().__init__()
        assert abs(quant_resi) > 1e-6, "quant_resi=0 (identity Phi) is not supported by the fused kernel"
        self.vocab_size, self.Cvae, self.using_znorm = vocab_size, Cvae, using_znorm
        self.v_patch_nums: Tuple[int, ...] = tuple(v_patch_nums)
        self.quant_resi_ratio = quant_resi
        if share_quant_resi == 1:
            self.quant_resi = _PhiBank(1, Cvae, quant_resi, "shared")
        elif share_quant_resi == 0:
            raise NotImplementedError("non-shared Phi (share_quant_resi=0) changes the state-dict layout; use 4 (default) or 1")
        else:
            self.quant_resi = _PhiBank(share_quant_resi, Cvae, quant_resi, "partial")
        self.register_buffer("ema_vocab_hit_SV", torch.zeros(len(self.v_patch_nums), vocab_size))
        self.beta = beta
        self.embedding = nn.Embedding(vocab_size, Cvae)
        self.prog_si = -1

    # ---- fused device path -------------------------------------------------------------------
    def _phi(self, si: int, SN: int) -> _PhiParams:
        return self.quant_resi[si / (SN - 1)]

    def next_input_from_idx(self, si: int, f_hat: torch.Tensor, idx_Bl: torch.Tensor, next_map: Optional[torch.Tensor] = None,
                            codebook: Optional[torch.Tensor] = None, f_rest: Optional[torch.Tensor] = None):
        """K5: f_hat += Phi(bicubic_up(codebook[idx])) in place; returns (f_hat, area_down(f_hat) to the next stage).
        Replaces models/var.py:205,210-211 + models/quant.py:187-196."""
        SN = len(self.v_patch_nums)
        B, pn, HW = idx_Bl.shape[0], self.v_patch_nums[si], self.v_patch_nums[-1]
        pn2 = self.v_patch_nums[si + 1] if si != SN - 1 else 0
        phi = self._phi(si, SN)
        if next_map is None and pn2:
            next_map = torch.empty(B, self.Cvae, pn2, pn2, device=f_hat.device, dtype=torch.float32)
        cb = self.embedding.weight if codebook is None else codebook
        _cabi.vq_next_input(idx_Bl, B, pn, HW, pn2, self.Cvae, cb, phi.weight, phi.bias, f_hat, next_map if pn2 else None,
                            resi_ratio=phi.resi_ratio, f_rest=f_rest)
        return f_hat, (next_map if pn2 else f_hat)

    def get_next_autoregressive_input(self, si: int, SN: int, f_hat: torch.Tensor, h_BChw: torch.Tensor):
        """Reference signature (models/quant.py:187): takes the already-embedded h (B,Cvae,pn,pn); mutates f_hat.
        The fused kernel gathers rows of a table, so h itself is passed as a (B*l, Cvae) table with identity indices."""
        B, Cv, pn, _ = h_BChw.shape
        assert SN == len(self.v_patch_nums) and pn == self.v_patch_nums[si]
        table = h_BChw.permute(0, 2, 3, 1).reshape(B * pn * pn, Cv).float().contiguous()
        ident = torch.arange(B * pn * pn, device=h_BChw.device, dtype=torch.int64).view(B, pn * pn)
        return self.next_input_from_idx(si, f_hat, ident, codebook=table)

    def embed_to_fhat(self, ms_h_BChw: List[torch.Tensor], all_to_max_scale=True, last_one=False):
        """models/quant.py:107-133 (all_to_max_scale=True only): accumulate every scale into f_hat."""
        assert all_to_max_scale, "all_to_max_scale=False is marked experimental in the reference and not supported"
        B, HW, SN = ms_h_BChw[0].shape[0], self.v_patch_nums[-1], len(self.v_patch_nums)
        f_hat = torch.zeros(B, self.Cvae, HW, HW, device=ms_h_BChw[0].device, dtype=torch.float32)
        outs = []
        for si in range(SN):
            self.get_next_autoregressive_input(si, SN, f_hat, ms_h_BChw[si])
            if not last_one:
                outs.append(f_hat.clone())
        return f_hat if last_one else outs

    def idxBl_to_var_input(self, gt_ms_idx_Bl: List[torch.Tensor]) -> Optional[torch.Tensor]:
        """Teacher-forcing input builder (models/quant.py:169-184): cat of area_down(f_hat) after each of the first
        SN-1 stages, as (B, L-first_l, Cvae)."""
        SN = len(self.v_patch_nums)
        B, HW = gt_ms_idx_Bl[0].shape[0], self.v_patch_nums[-1]
        f_hat = torch.zeros(B, self.Cvae, HW, HW, device=gt_ms_idx_Bl[0].device, dtype=torch.float32)
        nxt = []
        for si in range(SN - 1):
            _, nm = self.next_input_from_idx(si, f_hat, gt_ms_idx_Bl[si].contiguous())
            nxt.append(nm.view(B, self.Cvae, -1).transpose(1, 2))
        return torch.cat(nxt, dim=1) if nxt else None

    # ---- out of scope --------------------------------------------------------------------------
    def forward(self, *a, **k):
        raise NotImplementedError("VectorQuantizer2.forward is VAE training (SURVEY.md section 2 row 6): out of scope")

    # ---- encode side ------------------------------------------------------------
    def f_to_idxBl_or_fhat(self, f_BChw: torch.Tensor, to_fhat: bool, v_patch_nums=None):
        """Multi-scale residual quantisation of an encoder feature map (models/quant.py:135-166): per scale, area-downsample
        the residual, take the nearest codebook entry (``sdvar_vq_nearest_code``), add Phi(bicubic_up(embedding)) to f_hat
        (``sdvar_vq_next_input``), which also subtracts the same increment from the running residual ``f_rest`` exactly as
        the reference does (``f_hat.add_(h); f_rest.sub_(h)``, models/quant.py:162-163).  Returns the token lists
        (B, pn*pn) or the f_hat snapshots."""
        assert not self.using_znorm, "using_znorm=True (cosine nearest neighbour) is not supported"
        pns = tuple(v_patch_nums) if v_patch_nums is not None else self.v_patch_nums
        assert tuple(int(p if isinstance(p, int) else p[0]) for p in pns) == self.v_patch_nums, \
            "only the model's own patch_nums are supported"
        B, C, H, W = f_BChw.shape
        HW, SN = self.v_patch_nums[-1], len(self.v_patch_nums)
        assert C == self.Cvae and H == HW and W == HW, f"feature map {tuple(f_BChw.shape)} does not match the {HW}x{HW} latent grid"
        f_rest = f_BChw.detach().float().contiguous().clone()
        f_hat = torch.zeros_like(f_rest)
        cb = self.embedding.weight.detach().float().contiguous()
        out = []
        for si, pn in enumerate(self.v_patch_nums):
            z = torch.nn.functional.interpolate(f_rest, size=(pn, pn), mode="area") if si != SN - 1 else f_rest
            z_NC = z.permute(0, 2, 3, 1).reshape(-1, C).contiguous()
            idx_N = torch.empty(z_NC.shape[0], dtype=torch.int64, device=f_rest.device)
            _cabi.vq_nearest_code(z_NC, cb, z_NC.shape[0], C, self.vocab_size, idx_N)
            idx_Bl = idx_N.view(B, pn * pn)
            self.next_input_from_idx(si, f_hat, idx_Bl, codebook=cb, f_rest=f_rest)
            out.append(f_hat.clone() if to_fhat else idx_Bl)
        return out
