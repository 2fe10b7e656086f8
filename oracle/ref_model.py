"""PyTorch-CPU restatement of the reference's hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows (paths relative to the reference
repo).  The restatement consumes flat state dicts with the reference's own key names
(the checkpoint surface), so the same tensors can be loaded into the
reference (``load_state_dict(strict=True)``), this oracle and the CUDA engine.

Two arithmetic regimes:
  * ``mm='fp32'``  -- the reference's regime (fp32 parameters, models/var.py:125).
  * ``mm='bf16'``  -- emulates the CUDA engine's rounding points (GEMM operands rounded to
    bf16, fp32 accumulate, fp32 residual stream and logits) so that engine-vs-oracle
    comparisons are tight enough to catch indexing bugs hidden under bf16 noise.

Parity status: everything up to and including ``autoregressive_infer_cfg`` /
``forward_teacher`` / ``sd_test3`` is pinned against the real reference by
``oracle/make_golden.py`` and ``tests/test_oracle_vs_reference.py``.  ``verify_tokens`` and
``sd_generate`` restate the north-star speculative rule, which the reference does not
implement: PARITY UNPINNED.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

NEG_INF = float("-inf")


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def stage_table(patch_nums: Sequence[int]) -> Tuple[List[int], List[int], List[int]]:
    """l_s, begin_s, end_s of every stage (models/var.py:41-47)."""
    ls = [pn * pn for pn in patch_nums]
    ends = list(np.cumsum(ls))
    begins = [0] + ends[:-1]
    return ls, [int(b) for b in begins], [int(e) for e in ends]


def phi_index(si: int, SN: int, K: int = 4) -> int:
    """PhiPartiallyShared.__getitem__ (models/quant.py:218-226)."""
    ticks = np.linspace(1 / 3 / K, 1 - 1 / 3 / K, K) if K == 4 else np.linspace(1 / 2 / K, 1 - 1 / 2 / K, K)
    return int(np.argmin(np.abs(ticks - si / (SN - 1))).item())


def _r(x: torch.Tensor, mm: str) -> torch.Tensor:
    """round a GEMM operand the way the engine does"""
    return x.to(torch.bfloat16).float() if mm == "bf16" else x


def _linear(x, w, b, mm):
    y = F.linear(_r(x, mm), _r(w, mm))
    return y if b is None else y + b


# --------------------------------------------------------------------------------------
# sampler (models/helpers.py:6-19) and CFG mix (models/var.py:199-200)
# --------------------------------------------------------------------------------------
def cfg_mix(logits_2BlV: torch.Tensor, B: int, t: float) -> torch.Tensor:
    """(1+t)*cond - t*uncond, this exact association (models/var.py:199-200)."""
    return (1 + t) * logits_2BlV[:B] - t * logits_2BlV[B:]


def filter_top_k_top_p_(logits_BlV: torch.Tensor, top_k: int = 0, top_p: float = 0.0) -> torch.Tensor:
    """In-place top-k / top-p masking, helpers.py:8-15 verbatim in behaviour."""
    if top_k > 0:
        kth = logits_BlV.topk(top_k, largest=True, sorted=False, dim=-1)[0].amin(dim=-1, keepdim=True)
        logits_BlV.masked_fill_(logits_BlV < kth, NEG_INF)
    if top_p > 0:
        sorted_logits, sorted_idx = logits_BlV.sort(dim=-1, descending=False)
        rm = sorted_logits.softmax(dim=-1).cumsum_(dim=-1) <= (1 - top_p)
        rm[..., -1:] = False
        logits_BlV.masked_fill_(rm.scatter(sorted_idx.ndim - 1, sorted_idx, rm), NEG_INF)
    return logits_BlV


def sample_multinomial_(logits_BlV, top_k=0, top_p=0.0, rng=None) -> torch.Tensor:
    """helpers.py:6-19 with num_samples=1: returns idx (B,l); mutates logits."""
    B, l, V = logits_BlV.shape
    filter_top_k_top_p_(logits_BlV, top_k, top_p)
    return torch.multinomial(logits_BlV.softmax(dim=-1).view(-1, V), num_samples=1, replacement=True,
                             generator=rng).view(B, l)


def sample_with_noise_(logits_BlV, noise_NV, top_k=0, top_p=0.0) -> torch.Tensor:
    """Same as ``sample_multinomial_`` but with the Exp(1) noise pre-drawn:
    ATen's multinomial fast path for n_sample==1 is ``argmax(p / q)``, ``q ~ Exp(1)`` drawn as one
    (B*l, V) tensor (pin P4)."""
    B, l, V = logits_BlV.shape
    filter_top_k_top_p_(logits_BlV, top_k, top_p)
    p = logits_BlV.softmax(dim=-1).view(-1, V)
    return torch.argmax(p / noise_NV.view(-1, V), dim=-1).view(B, l)


# --------------------------------------------------------------------------------------
# speculative verify (north_star item 3) -- PARITY UNPINNED
# --------------------------------------------------------------------------------------
def verify_tokens(xt_NV, xd_NV, draft_idx_N, u_N, noise_NV):
    """Per token: p=softmax(xt), q=softmax(xd); accept iff u*q[d] < p[d]; on reject
    out = argmax(max(0,p-q)/noise) (if the residual is identically 0: argmax(p/noise)).
    Returns (out_idx int64 (N,), accept bool (N,), p_d, q_d)."""
    # exp * (1/Z), the association the C spec / kernel use (one division per row)
    et = (xt_NV - xt_NV.amax(dim=-1, keepdim=True)).exp()
    ed = (xd_NV - xd_NV.amax(dim=-1, keepdim=True)).exp()
    p = et * (1.0 / et.sum(dim=-1, keepdim=True))
    q = ed * (1.0 / ed.sum(dim=-1, keepdim=True))
    ar = torch.arange(p.shape[0])
    p_d, q_d = p[ar, draft_idx_N], q[ar, draft_idx_N]
    accept = (u_N * q_d) < p_d
    r = (p - q).clamp_min(0.0)
    res = torch.where(r.sum(-1, keepdim=True) > 0, r, p)
    repaired = torch.argmax(res / noise_NV, dim=-1)
    out = torch.where(accept, draft_idx_N, repaired)
    return out, accept, p_d, q_d


def first_reject_scan(accept_Bl: torch.Tensor):
    """per (image, stage): index of first rejected token (l if none) and #accepted tokens."""
    B, l = accept_Bl.shape
    pos = torch.arange(l).expand(B, l)
    first = torch.where(accept_Bl, torch.full_like(pos, l), pos).amin(dim=1)
    return first, accept_Bl.sum(dim=1)


# --------------------------------------------------------------------------------------
# VQ next-input (models/quant.py:187-206, 218-226)
# --------------------------------------------------------------------------------------
class RefVQ:
    def __init__(self, sd: Dict[str, torch.Tensor], patch_nums: Sequence[int], prefix: str = "quantize."):
        self.patch_nums = tuple(patch_nums)
        self.codebook = sd[prefix + "embedding.weight"].float()
        self.V, self.Cvae = self.codebook.shape
        self.phi_w, self.phi_b = [], []
        i = 0
        while f"{prefix}quant_resi.qresi_ls.{i}.weight" in sd:
            self.phi_w.append(sd[f"{prefix}quant_resi.qresi_ls.{i}.weight"].float())
            self.phi_b.append(sd[f"{prefix}quant_resi.qresi_ls.{i}.bias"].float())
            i += 1
        self.resi_ratio = 0.5  # quant_resi=0.5 (models/vqvae.py:22)

    def embedding(self, idx_Bl):
        return self.codebook[idx_Bl]  # (B,l,Cvae)  models/quant.py:39

    def phi(self, si: int, h: torch.Tensor) -> torch.Tensor:
        """Phi.forward (models/quant.py:205-206): h*(1-r) + conv3x3(h)*r"""
        k = phi_index(si, len(self.patch_nums), len(self.phi_w))
        return h.mul(1 - self.resi_ratio) + F.conv2d(h, self.phi_w[k], self.phi_b[k], padding=1).mul_(self.resi_ratio)

    def next_input(self, si: int, f_hat: torch.Tensor, idx_Bl: torch.Tensor):
        """embedding -> transpose/reshape (models/var.py:205,210) -> get_next_autoregressive_input
        (models/quant.py:187-196).  Mutates and returns f_hat; next_map is (B,Cvae,pn',pn')."""
        SN = len(self.patch_nums)
        B = idx_Bl.shape[0]
        pn = self.patch_nums[si]
        HW = self.patch_nums[-1]
        h = self.embedding(idx_Bl).transpose(1, 2).reshape(B, self.Cvae, pn, pn)
        if si != SN - 1:
            f_hat.add_(self.phi(si, F.interpolate(h, size=(HW, HW), mode="bicubic")))
            pn2 = self.patch_nums[si + 1]
            return f_hat, F.interpolate(f_hat, size=(pn2, pn2), mode="area")
        f_hat.add_(self.phi(si, h))
        return f_hat, f_hat

    def f_to_idxBl_or_fhat(self, f_BChw: torch.Tensor, to_fhat: bool, nearest=None):
        """Multi-scale residual quantisation, reference models/quant.py:135-166 (using_znorm=False).  `nearest(z_NC) -> idx_N`
        overrides the reference's addmm/argmin (default) -- tests pass the C spec's fixed-order search."""
        B, C, H, W = f_BChw.shape
        SN = len(self.patch_nums)
        f_rest = f_BChw.detach().float().clone()
        f_hat = torch.zeros_like(f_rest)
        out = []
        for si, pn in enumerate(self.patch_nums):
            z = F.interpolate(f_rest, size=(pn, pn), mode="area") if si != SN - 1 else f_rest
            z_NC = z.permute(0, 2, 3, 1).reshape(-1, C)
            if nearest is None:
                d = torch.sum(z_NC.square(), dim=1, keepdim=True) + torch.sum(self.codebook.square(), dim=1, keepdim=False)
                d.addmm_(z_NC, self.codebook.T, alpha=-2, beta=1)
                idx_N = torch.argmin(d, dim=1)
            else:
                idx_N = nearest(z_NC.contiguous())
            h = self.embedding(idx_N.view(B, pn, pn)).permute(0, 3, 1, 2)
            h = F.interpolate(h, size=(H, W), mode="bicubic").contiguous() if si != SN - 1 else h.contiguous()
            h = self.phi(si, h)
            f_hat.add_(h)
            f_rest.sub_(h)
            out.append(f_hat.clone() if to_fhat else idx_N.reshape(B, pn * pn))
        return out

    # -- closed forms the CUDA kernel implements (pin P5) --
    @staticmethod
    def bicubic_matrix(pn: int, HW: int) -> torch.Tensor:
        """(HW, pn) interpolation matrix of F.interpolate(mode='bicubic', align_corners=False):
        Keys cubic A=-0.75, source coord (o+0.5)*pn/HW-0.5, taps clamped to [0,pn-1]."""
        A = -0.75
        M = torch.zeros(HW, pn, dtype=torch.float64)
        scale = pn / HW
        for o in range(HW):
            sx = (o + 0.5) * scale - 0.5
            ix = math.floor(sx)
            t = sx - ix
            w = [((A * (t + 1) - 5 * A) * (t + 1) + 8 * A) * (t + 1) - 4 * A,
                 ((A + 2) * t - (A + 3)) * t * t + 1,
                 ((A + 2) * (1 - t) - (A + 3)) * (1 - t) * (1 - t) + 1,
                 ((A * (2 - t) - 5 * A) * (2 - t) + 8 * A) * (2 - t) - 4 * A]
            for k in range(4):
                M[o, min(max(ix - 1 + k, 0), pn - 1)] += w[k]
        return M

    @staticmethod
    def area_matrix(HW: int, pn2: int) -> torch.Tensor:
        """(pn2, HW) matrix of F.interpolate(mode='area') == adaptive_avg_pool:
        window [floor(o*HW/pn2), ceil((o+1)*HW/pn2))."""
        M = torch.zeros(pn2, HW, dtype=torch.float64)
        for o in range(pn2):
            a, b = (o * HW) // pn2, -((-(o + 1) * HW) // pn2)
            M[o, a:b] = 1.0 / (b - a)
        return M

    def next_input_closed(self, si: int, f_hat: torch.Tensor, idx_Bl: torch.Tensor):
        SN = len(self.patch_nums)
        B = idx_Bl.shape[0]
        pn, HW = self.patch_nums[si], self.patch_nums[-1]
        h = self.embedding(idx_Bl).transpose(1, 2).reshape(B, self.Cvae, pn, pn).double()
        if si != SN - 1:
            Mb = self.bicubic_matrix(pn, HW)
            h = torch.einsum("yp,bcpq,xq->bcyx", Mb, h, Mb)
        k = phi_index(si, SN, len(self.phi_w))
        hh = 0.5 * h + 0.5 * F.conv2d(h, self.phi_w[k].double(), self.phi_b[k].double(), padding=1)
        f_hat.add_(hh.float())
        if si != SN - 1:
            Ma = self.area_matrix(HW, self.patch_nums[si + 1])
            return f_hat, torch.einsum("oy,bcyx,px->bcop", Ma, f_hat.double(), Ma).float()
        return f_hat, f_hat


# --------------------------------------------------------------------------------------
# VQVAE decoder (models/vqvae.py:62-63, models/basic_vae.py:163-226)
# --------------------------------------------------------------------------------------
class RefDecoder:
    def __init__(self, sd: Dict[str, torch.Tensor], ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2):
        self.sd = {k: v.float() for k, v in sd.items() if k.startswith(("decoder.", "post_quant_conv."))}
        self.ch_mult, self.nrb = ch_mult, num_res_blocks

    def _conv(self, name, x, pad):
        return F.conv2d(x, self.sd[name + ".weight"], self.sd[name + ".bias"], padding=pad)

    def _gn(self, name, x):
        return F.group_norm(x, 32, self.sd[name + ".weight"], self.sd[name + ".bias"], eps=1e-6)

    def _res(self, name, x):  # basic_vae.py:57-60
        h = self._conv(name + ".conv1", F.silu(self._gn(name + ".norm1", x)), 1)
        h = self._conv(name + ".conv2", F.silu(self._gn(name + ".norm2", h)), 1)
        sc = self._conv(name + ".nin_shortcut", x, 0) if (name + ".nin_shortcut.weight") in self.sd else x
        return sc + h

    def _attn(self, name, x):  # basic_vae.py:74-94
        qkv = self._conv(name + ".qkv", self._gn(name + ".norm", x), 0)
        B, C3, H, W = qkv.shape
        C = C3 // 3
        q, k, v = qkv.reshape(B, 3, C, H * W).unbind(1)
        w = torch.bmm(q.permute(0, 2, 1), k).mul_(C ** -0.5).softmax(dim=2)
        h = torch.bmm(v, w.permute(0, 2, 1)).view(B, C, H, W)
        return x + self._conv(name + ".proj_out", h, 0)

    def fhat_to_img(self, f_hat: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """``taps`` (optional dict) receives the feature map after conv_in, after the mid block and after every up level
        (keys 'conv_in', 'mid', 'up.4' ... 'up.0'): per-layer parity targets for the device decoder."""
        def tap(k, v):
            if taps is not None:
                taps[k] = v.clone()
        h = self._conv("post_quant_conv", f_hat.float(), 1)
        h = self._conv("decoder.conv_in", h, 1)
        tap("conv_in", h)
        h = self._res("decoder.mid.block_2", self._attn("decoder.mid.attn_1", self._res("decoder.mid.block_1", h)))
        tap("mid", h)
        nres = len(self.ch_mult)
        for i_level in reversed(range(nres)):
            for i_block in range(self.nrb + 1):
                h = self._res(f"decoder.up.{i_level}.block.{i_block}", h)
                if i_level == nres - 1:
                    h = self._attn(f"decoder.up.{i_level}.attn.{i_block}", h)
            if i_level != 0:
                h = self._conv(f"decoder.up.{i_level}.upsample.conv", F.interpolate(h, scale_factor=2, mode="nearest"), 1)
            tap(f"up.{i_level}", h)
        h = self._conv("decoder.conv_out", F.silu(self._gn("decoder.norm_out", h)), 1)
        return h.clamp_(-1, 1)


class RefEncoder(RefDecoder):
    """quant_conv(Encoder(img)), reference models/vqvae.py:66 + models/basic_vae.py:99-160 (encode side)."""

    def __init__(self, sd: Dict[str, torch.Tensor], ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2):
        self.sd = {k: v.float() for k, v in sd.items() if k.startswith(("encoder.", "quant_conv."))}
        self.ch_mult, self.nrb = ch_mult, num_res_blocks

    def encode_features(self, img: torch.Tensor) -> torch.Tensor:
        h = self._conv("encoder.conv_in", img.float(), 1)
        nres = len(self.ch_mult)
        for i_level in range(nres):
            for i_block in range(self.nrb):
                h = self._res(f"encoder.down.{i_level}.block.{i_block}", h)
                if i_level == nres - 1:
                    h = self._attn(f"encoder.down.{i_level}.attn.{i_block}", h)
            if i_level != nres - 1:   # Downsample2x, basic_vae.py:36-37
                n = f"encoder.down.{i_level}.downsample.conv"
                h = F.conv2d(F.pad(h, pad=(0, 1, 0, 1), mode="constant", value=0), self.sd[n + ".weight"], self.sd[n + ".bias"], stride=2)
        h = self._res("encoder.mid.block_2", self._attn("encoder.mid.attn_1", self._res("encoder.mid.block_1", h)))
        h = self._conv("encoder.conv_out", F.silu(self._gn("encoder.norm_out", h)), 1)
        return self._conv("quant_conv", h, 1)


# --------------------------------------------------------------------------------------
# VAR transformer (models/var.py:22-259, models/basic_var.py)
# --------------------------------------------------------------------------------------
class RefVAR:
    def __init__(self, sd: Dict[str, torch.Tensor], patch_nums: Sequence[int], num_heads: Optional[int] = None,
                 num_classes: int = 1000, attn_l2_norm: bool = True, mm: str = "fp32", norm_eps: float = 1e-6):
        self.sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
        self.patch_nums = tuple(patch_nums)
        self.ls, self.begins, self.ends = stage_table(patch_nums)
        self.L = self.ends[-1]
        self.C = self.sd["pos_1LC"].shape[-1]
        self.Cvae = self.sd["word_embed.weight"].shape[1]
        self.V = self.sd["head.weight"].shape[0]
        self.depth = 0
        while f"blocks.{self.depth}.attn.mat_qkv.weight" in self.sd:
            self.depth += 1
        self.H = num_heads or self.depth
        self.hd = self.C // self.H
        self.num_classes = num_classes
        self.shared_aln = "shared_ada_lin.1.weight" in self.sd
        self.l2 = attn_l2_norm
        self.mm = mm
        self.eps = norm_eps
        self.first_l = self.ls[0]
        self.kv: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None] * self.depth
        self.caching = False
        self.lvl_pos = self.sd["lvl_embed.weight"][self.sd["lvl_1L"]] + self.sd["pos_1LC"]  # var.py:164

    # ---- KV cache (basic_var.py:87,107-109) + rollback-by-length (new spec, fixes D4) ----
    def kv_caching(self, enable: bool):
        self.caching = enable
        self.kv = [None] * self.depth

    def kv_len(self) -> int:
        return 0 if self.kv[0] is None else self.kv[0][0].shape[2]

    def kv_truncate(self, n: int):
        self.kv = [None if (kv is None or n == 0) else (kv[0][:, :, :n], kv[1][:, :, :n]) for kv in self.kv]

    # ---- conditioning ----
    def cond(self, label_B: torch.Tensor) -> torch.Tensor:
        """class_emb(cat(label, num_classes)) (2B,C)  var.py:162"""
        return self.sd["class_emb.weight"][torch.cat((label_B, torch.full_like(label_B, self.num_classes)), dim=0)]

    def _ada6(self, i: int, cond_BD: torch.Tensor) -> torch.Tensor:
        """(2B,1,6,C) gamma1,gamma2,scale1,scale2,shift1,shift2  basic_var.py:152-156, var.py:16-19,81"""
        if self.shared_aln:
            g = _linear(F.silu(cond_BD), self.sd["shared_ada_lin.1.weight"], self.sd["shared_ada_lin.1.bias"], self.mm)
            return self.sd[f"blocks.{i}.ada_gss"] + g.view(-1, 1, 6, self.C)
        return _linear(F.silu(cond_BD), self.sd[f"blocks.{i}.ada_lin.1.weight"], self.sd[f"blocks.{i}.ada_lin.1.bias"],
                       self.mm).view(-1, 1, 6, self.C)

    # ---- one block (basic_var.py:90-119, 44-52, 152-159) ----
    def _attn(self, i: int, x: torch.Tensor, attn_bias: Optional[torch.Tensor]) -> torch.Tensor:
        p = f"blocks.{i}.attn."
        Bx, L, C = x.shape
        bias = torch.cat((self.sd[p + "q_bias"], self.sd[p + "zero_k_bias"], self.sd[p + "v_bias"]))
        qkv = _linear(x, self.sd[p + "mat_qkv.weight"], bias, self.mm).view(Bx, L, 3, self.H, self.hd)
        q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(dim=0)  # BHLc
        if self.l2:
            scale_mul = self.sd[p + "scale_mul_1H11"].clamp_max(math.log(100)).exp()
            q = F.normalize(q, dim=-1).mul(scale_mul)
            k = F.normalize(k, dim=-1)
            scale = 1.0
        else:
            scale = 0.25 / math.sqrt(self.hd)
        q, k, v = _r(q, self.mm), _r(k, self.mm), _r(v, self.mm)
        if self.caching:
            if self.kv[i] is not None:
                k = torch.cat((self.kv[i][0], k), dim=2)
                v = torch.cat((self.kv[i][1], v), dim=2)
            self.kv[i] = (k, v)
        if self.mm == "bf16":
            s = (q @ k.transpose(-2, -1)) * scale
            if attn_bias is not None:
                s = s + attn_bias
            pr = s.softmax(dim=-1)
            # engine: P is rounded to bf16 before P@V, the row sum is taken over the rounded P
            e = torch.exp(s - s.amax(dim=-1, keepdim=True))
            eb = e.to(torch.bfloat16).float()
            o = (eb @ v) / eb.sum(dim=-1, keepdim=True)
            del pr
        else:
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_bias, scale=scale)
        o = _r(o.transpose(1, 2).reshape(Bx, L, C), self.mm)
        return _linear(o, self.sd[p + "proj.weight"], self.sd[p + "proj.bias"], self.mm)

    def _ffn(self, i: int, x: torch.Tensor) -> torch.Tensor:
        p = f"blocks.{i}.ffn."
        h = F.gelu(_linear(x, self.sd[p + "fc1.weight"], self.sd[p + "fc1.bias"], self.mm), approximate="tanh")
        return _linear(h, self.sd[p + "fc2.weight"], self.sd[p + "fc2.bias"], self.mm)

    def block(self, i: int, x: torch.Tensor, cond_BD: torch.Tensor, attn_bias) -> torch.Tensor:
        g1, g2, s1, s2, b1, b2 = self._ada6(i, cond_BD).unbind(2)
        ln = lambda t: F.layer_norm(t, (self.C,), eps=self.eps)
        x = x + self._attn(i, ln(x).mul(s1.add(1)).add_(b1), attn_bias).mul_(g1)
        x = x + self._ffn(i, ln(x).mul(s2.add(1)).add_(b2)).mul(g2)
        return x

    def blocks(self, x: torch.Tensor, cond_BD: torch.Tensor, attn_bias=None) -> torch.Tensor:
        for i in range(self.depth):
            x = self.block(i, x, cond_BD, attn_bias)
        return x

    def get_logits(self, h: torch.Tensor, cond_BD: torch.Tensor) -> torch.Tensor:
        """var.py:119-125 + AdaLNBeforeHead basic_var.py:172-174 (scale, shift order)."""
        a = _linear(F.silu(cond_BD), self.sd["head_nm.ada_lin.1.weight"], self.sd["head_nm.ada_lin.1.bias"], self.mm)
        scale, shift = a.view(-1, 1, 2, self.C).unbind(2)
        hn = F.layer_norm(h.float(), (self.C,), eps=self.eps).mul(scale.add(1)).add_(shift)
        return _linear(hn, self.sd["head.weight"], self.sd["head.bias"], self.mm).float()

    # ---- stage inputs (var.py:178-188) ----
    def first_map(self, cond_BD: torch.Tensor) -> torch.Tensor:
        n = cond_BD.shape[0]
        return (cond_BD.unsqueeze(1).expand(n, self.first_l, -1) + self.sd["pos_start"].expand(n, self.first_l, -1)
                + self.lvl_pos[:, :self.first_l])

    def embed_map(self, si: int, next_map_BChw: torch.Tensor) -> torch.Tensor:
        """word_embed(next_map.view(B,Cvae,-1).T) + lvl_pos[begin:end], then repeat(2,1,1)  var.py:185-188"""
        B = next_map_BChw.shape[0]
        t = next_map_BChw.reshape(B, self.Cvae, -1).transpose(1, 2)
        x = F.linear(t, self.sd["word_embed.weight"], self.sd["word_embed.bias"]) + self.lvl_pos[:, self.begins[si]:self.ends[si]]
        return x.repeat(2, 1, 1)

    # ---- baseline loop (var.py:128-215) ----
    @torch.no_grad()
    def autoregressive_infer_cfg(self, vq: RefVQ, B: int, label_B: torch.Tensor, cfg=1.5, top_k=0, top_p=0.0,
                                 rng: Optional[torch.Generator] = None, noise: Optional[List[torch.Tensor]] = None,
                                 record: Optional[dict] = None):
        """Returns (f_hat, [idx_Bl per stage]).  ``noise`` (list of (B*l,V) Exp(1) tensors) replaces the
        generator when given.  ``record`` collects per-stage tensors for golden dumps."""
        K = len(self.patch_nums)
        cond_BD = self.cond(label_B)
        f_hat = cond_BD.new_zeros(B, self.Cvae, self.patch_nums[-1], self.patch_nums[-1])
        self.kv_caching(True)
        idxs = []
        next_map = None
        for si in range(K):
            x = self.first_map(cond_BD) if si == 0 else self.embed_map(si, next_map)
            h = self.blocks(x, cond_BD, None)
            logits = self.get_logits(h, cond_BD)
            t = cfg * (si / (K - 1))
            mixed = cfg_mix(logits, B, t)
            if record is not None:
                record.setdefault("x", []).append(x.clone()); record.setdefault("h", []).append(h.clone())
                record.setdefault("logits", []).append(logits.clone()); record.setdefault("mixed", []).append(mixed.clone())
            if noise is not None:
                idx = sample_with_noise_(mixed, noise[si], top_k, top_p)
            else:
                idx = sample_multinomial_(mixed, top_k, top_p, rng)
            idxs.append(idx)
            f_hat, next_map = vq.next_input(si, f_hat, idx)
            if record is not None:
                record.setdefault("f_hat", []).append(f_hat.clone()); record.setdefault("next_map", []).append(next_map.clone())
        self.kv_caching(False)
        return f_hat, idxs

    # ---- teacher-forced forward (var.py:217-259, mask var.py:108-113), cond_drop disabled ----
    @torch.no_grad()
    def forward_teacher(self, label_B: torch.Tensor, x_BLCv_wo_first_l: torch.Tensor) -> torch.Tensor:
        B = x_BLCv_wo_first_l.shape[0]
        cond_BD = self.sd["class_emb.weight"][label_B]
        sos = cond_BD.unsqueeze(1).expand(B, self.first_l, -1) + self.sd["pos_start"].expand(B, self.first_l, -1)
        x = torch.cat((sos, F.linear(x_BLCv_wo_first_l.float(), self.sd["word_embed.weight"], self.sd["word_embed.bias"])), dim=1)
        x = x + self.lvl_pos
        h = self.blocks(x, cond_BD, self.sd["attn_bias_for_masking"])
        return self.get_logits(h, cond_BD)

    # ---- multi-stage window pass = the shape of "verify" (new spec; fixes D1-D6) ----
    def window_mask(self, s0: int, g: int) -> torch.Tensor:
        """Additive mask rows = tokens of stages [s0,s0+g), cols = keys [0, end of stage s0+g-1):
        query of stage j sees keys < ends[j] (block-causal, var.py:108-113 restricted to the window)."""
        q_lvl = self.sd["lvl_1L"][0, self.begins[s0]:self.ends[s0 + g - 1]].view(-1, 1)
        k_lvl = self.sd["lvl_1L"][0, :self.ends[s0 + g - 1]].view(1, -1)
        return torch.where(q_lvl >= k_lvl, 0.0, NEG_INF).view(1, 1, q_lvl.shape[0], k_lvl.shape[1])

    def forward_window(self, s0: int, xs: List[torch.Tensor], cond_BD: torch.Tensor) -> List[torch.Tensor]:
        """One pass over the concatenated inputs of stages s0..s0+g-1 (each (2B,l_s,C)) with the KV
        cache holding stages < s0.  Returns per-stage raw logits (2B,l_s,V)."""
        g = len(xs)
        assert self.caching and self.kv_len() == self.begins[s0]
        x = torch.cat(xs, dim=1)
        mask = None if g == 1 else self.window_mask(s0, g)
        logits = self.get_logits(self.blocks(x, cond_BD, mask), cond_BD)
        out, o = [], 0
        for j in range(g):
            out.append(logits[:, o:o + self.ls[s0 + j]])
            o += self.ls[s0 + j]
        return out


# --------------------------------------------------------------------------------------
# hand-over variant (models/var.py:605-865, sd_mask=0) -- pins P1/P2
# --------------------------------------------------------------------------------------
@torch.no_grad()
def sd_test3(draft: RefVAR, target: RefVAR, vq: RefVQ, B: int, label_B: torch.Tensor, cfg=1.5, top_k=0, top_p=0.0,
             entry_num: int = 10, rng: Optional[torch.Generator] = None, noise: Optional[List[torch.Tensor]] = None):
    """Draft runs stages [0,entry_num), target runs [entry_num,K) starting from the draft's f_hat, being fed
    only the current stage's tokens with an EMPTY KV cache (var.py:817-825): it never attends to the
    drafted prefix.  One generator for both models (var.py:641-644)."""
    K = len(draft.patch_nums)
    idxs = []

    def samp(mixed, si):
        return sample_with_noise_(mixed, noise[si], top_k, top_p) if noise is not None else sample_multinomial_(mixed, top_k, top_p, rng)

    cond_d = draft.cond(label_B)
    f_hat = cond_d.new_zeros(B, draft.Cvae, draft.patch_nums[-1], draft.patch_nums[-1])
    draft.kv_caching(True)
    next_map = None
    for si in range(min(entry_num, K)):
        x = draft.first_map(cond_d) if si == 0 else draft.embed_map(si, next_map)
        mixed = cfg_mix(draft.get_logits(draft.blocks(x, cond_d, None), cond_d), B, cfg * si / (K - 1))
        idxs.append(samp(mixed, si))
        f_hat, next_map = vq.next_input(si, f_hat, idxs[-1])
    draft.kv_caching(False)
    if entry_num >= K:
        return f_hat, idxs
    cond_t = target.cond(label_B)
    target.kv_caching(True)
    for si in range(entry_num, K):
        x = target.first_map(cond_t) if si == 0 else target.embed_map(si, next_map)
        mixed = cfg_mix(target.get_logits(target.blocks(x, cond_t, None), cond_t), B, cfg * si / (K - 1))
        idxs.append(samp(mixed, si))
        f_hat, next_map = vq.next_input(si, f_hat, idxs[-1])
    target.kv_caching(False)
    return f_hat, idxs


# --------------------------------------------------------------------------------------
# the draft -> verify loop, corrected restatement of models/var.py:1285-1383 (PARITY UNPINNED)
# --------------------------------------------------------------------------------------
class ReplayNoise:
    """Noise provider shared by oracle and engine in parity tests: draws on CPU from seeded
    generators in the order the loop spec fixes ('noise streams')."""

    def __init__(self, seed: int, device="cpu"):
        self.g = {k: torch.Generator().manual_seed(seed * 4 + i) for i, k in enumerate(("draft", "target", "u", "resample"))}
        self.device = device

    def exponential(self, stream: str, rows: int, V: int) -> torch.Tensor:
        return torch.empty(rows, V).exponential_(generator=self.g[stream]).to(self.device)

    def uniform(self, stream: str, rows: int) -> torch.Tensor:
        return torch.rand(rows, generator=self.g[stream]).to(self.device)


@torch.no_grad()
def sd_generate(draft: RefVAR, target: RefVAR, vq: RefVQ, B: int, label_B: torch.Tensor, noise, cfg=1.5, gamma=2,
                top_k=0, top_p=0.0, accept_rule: str = "speculative", match_threshold: float = 0.5, verify_mode: str = "window"):
    """Spec of ``sdvar_autoregressive_infer_cfg_parallel_v1`` with defects D1-D11 resolved:

    round at stage s, window g=min(gamma,K-s):
      1. draft runs stages s..s+g-1 incrementally (var.py:949-1024), each stage's input being
         area_down(f_hat) of the previous drafted stage (D2), stage 0 being the sos map (D3);
      2. target runs ONE block-causal pass over the same g stage inputs (var.py:1026-1070 intent) on top of
         its KV cache of accepted stages (D5,D6);
      3. per token ``verify_tokens`` (speculative rule) or top-1 match (reference rule, var.py:1199-1222);
      4. lock-step advance a = 1 + #leading window stages with no reject in ANY image (capped at g): stages
         s..s+a-2 keep the draft tokens, stage s+a-1 keeps accepted tokens and the target's repairs (D8);
      5. both KV caches are truncated to ends[s+a-1] (D4), f_hat is rebuilt from the snapshot after stage
         s+a-2 plus the final tokens of stage s+a-1 (D7).
    ``verify_mode='lazy'`` restates the engine's stage-by-stage schedule: stage s+j is drafted and run through
    the target (one single-stage pass on top of the cache) only if every stage before it in the window was accepted whole;
    the noise of the stages never reached is drawn and dropped so the three streams stay where the window schedule leaves
    them.  Same tokens, accept flags and advances as 'window'; ``target_passes`` / ``draft_stages`` count the work actually done.
    Returns (f_hat, final idx per stage, stats dict)."""
    K = len(draft.patch_nums)
    V = draft.V
    if verify_mode == "lazy":
        assert accept_rule == "speculative"
        return _sd_generate_lazy(draft, target, vq, B, label_B, noise, cfg, gamma, top_k, top_p)
    cond_d, cond_t = draft.cond(label_B), target.cond(label_B)
    f_hat = cond_d.new_zeros(B, draft.Cvae, draft.patch_nums[-1], draft.patch_nums[-1])
    draft.kv_caching(True); target.kv_caching(True)
    s, next_map = 0, None
    final_idx: List[torch.Tensor] = []
    stats = dict(rounds=0, target_passes=0, draft_stages=0, accepted_tokens=0, rejected_tokens=0,
                 advance=[], stage_accept_tokens=[0] * K, stage_tokens=[0] * K)
    while s < K:
        g = min(gamma, K - s)
        fh = f_hat.clone()
        maps, snaps, idx_d, mixed_d = [next_map], [], [], []
        for j in range(g):
            si = s + j
            x = draft.first_map(cond_d) if si == 0 else draft.embed_map(si, maps[j])
            mixed = cfg_mix(draft.get_logits(draft.blocks(x, cond_d, None), cond_d), B, cfg * si / (K - 1))
            n_d = noise.exponential("draft", B * draft.ls[si], V)
            idx = sample_with_noise_(mixed, n_d, top_k, top_p)  # mixed is masked in place
            fh, nm = vq.next_input(si, fh, idx)
            idx_d.append(idx); mixed_d.append(mixed); snaps.append(fh.clone()); maps.append(nm)
            stats["draft_stages"] += 1
        xs = [target.first_map(cond_t) if s + j == 0 else target.embed_map(s + j, maps[j]) for j in range(g)]
        logits_t = target.forward_window(s, xs, cond_t)
        stats["target_passes"] += 1
        out_idx, stage_ok = [], []
        for j in range(g):
            si = s + j
            l = draft.ls[si]
            mixed_t = filter_top_k_top_p_(cfg_mix(logits_t[j], B, cfg * si / (K - 1)), top_k, top_p)
            u = noise.uniform("u", B * l)
            n_r = noise.exponential("resample", B * l, V)
            if accept_rule == "speculative":
                o, acc, _, _ = verify_tokens(mixed_t.view(-1, V), mixed_d[j].view(-1, V), idx_d[j].view(-1), u, n_r)
                out_idx.append(o.view(B, l))
                acc = acc.view(B, l)
                stage_ok.append(bool(acc.all()))
            else:  # reference rule: top-1 match rate over the whole local batch >= threshold (var.py:1199-1217)
                match = (mixed_t.argmax(dim=-1) == idx_d[j])
                ok = bool(match.float().mean().item() >= match_threshold)
                stage_ok.append(ok)
                acc = match if ok else torch.zeros_like(match)
                if ok:
                    out_idx.append(idx_d[j])
                else:  # repair: the target samples the stage itself (replaces the reference's break, D8)
                    p = mixed_t.softmax(dim=-1).view(-1, V)
                    out_idx.append(torch.argmax(p / n_r, dim=-1).view(B, l))
            if j == 0 or all(stage_ok[:j]):
                stats["stage_accept_tokens"][si] += int(acc.sum()); stats["stage_tokens"][si] += B * l
        n_ok = 0
        while n_ok < g and stage_ok[n_ok]:
            n_ok += 1
        a = min(n_ok + 1, g)
        for j in range(a - 1):
            final_idx.append(idx_d[j])
        final_idx.append(out_idx[a - 1])
        base = snaps[a - 2] if a >= 2 else f_hat
        f_hat, next_map = vq.next_input(s + a - 1, base.clone(), out_idx[a - 1])
        draft.kv_truncate(draft.ends[s + a - 1]); target.kv_truncate(target.ends[s + a - 1])
        stats["rounds"] += 1; stats["advance"].append(a)
        s += a
    draft.kv_caching(False); target.kv_caching(False)
    stats["accepted_tokens"] = sum(stats["stage_accept_tokens"])
    stats["rejected_tokens"] = sum(stats["stage_tokens"]) - stats["accepted_tokens"]
    return f_hat, final_idx, stats


@torch.no_grad()
def _sd_generate_lazy(draft: RefVAR, target: RefVAR, vq: RefVQ, B, label_B, noise, cfg, gamma, top_k, top_p):
    K, V = len(draft.patch_nums), draft.V
    cond_d, cond_t = draft.cond(label_B), target.cond(label_B)
    f_hat = cond_d.new_zeros(B, draft.Cvae, draft.patch_nums[-1], draft.patch_nums[-1])
    draft.kv_caching(True); target.kv_caching(True)
    s, next_map = 0, None
    final_idx: List[torch.Tensor] = []
    stats = dict(rounds=0, target_passes=0, draft_stages=0, accepted_tokens=0, rejected_tokens=0,
                 advance=[], stage_accept_tokens=[0] * K, stage_tokens=[0] * K)
    while s < K:
        g = min(gamma, K - s)
        fh = f_hat.clone()
        maps, snaps, idx_d, out_idx = [next_map], [], [], []
        u_all = [None] * g
        n_ok = 0
        # the u / resample draws of the whole window come first in their streams, exactly as in the window schedule
        for j in range(g):
            l = draft.ls[s + j]
            u_all[j] = (noise.uniform("u", B * l), noise.exponential("resample", B * l, V))
        for j in range(g):
            si = s + j
            l = draft.ls[si]
            x = draft.first_map(cond_d) if si == 0 else draft.embed_map(si, maps[j])
            mixed_d = cfg_mix(draft.get_logits(draft.blocks(x, cond_d, None), cond_d), B, cfg * si / (K - 1))
            idx = sample_with_noise_(mixed_d, noise.exponential("draft", B * l, V), top_k, top_p)
            fh, nm = vq.next_input(si, fh, idx)
            idx_d.append(idx); snaps.append(fh.clone()); maps.append(nm)
            stats["draft_stages"] += 1
            xt = target.first_map(cond_t) if si == 0 else target.embed_map(si, maps[j])
            logits_t = target.forward_window(si, [xt], cond_t)[0]
            stats["target_passes"] += 1
            mixed_t = filter_top_k_top_p_(cfg_mix(logits_t, B, cfg * si / (K - 1)), top_k, top_p)
            o, acc, _, _ = verify_tokens(mixed_t.view(-1, V), mixed_d.view(-1, V), idx.view(-1), u_all[j][0], u_all[j][1])
            out_idx.append(o.view(B, l))
            stats["stage_accept_tokens"][si] += int(acc.sum()); stats["stage_tokens"][si] += B * l
            if not bool(acc.all()):
                break
            n_ok += 1
        for j in range(len(idx_d), g):           # never drafted: keep the 'draft' stream in step
            noise.exponential("draft", B * draft.ls[s + j], V)
        a = min(n_ok + 1, g)
        for j in range(a - 1):
            final_idx.append(idx_d[j])
        final_idx.append(out_idx[a - 1])
        base = snaps[a - 2] if a >= 2 else f_hat
        f_hat, next_map = vq.next_input(s + a - 1, base.clone(), out_idx[a - 1])
        draft.kv_truncate(draft.ends[s + a - 1]); target.kv_truncate(target.ends[s + a - 1])
        stats["rounds"] += 1; stats["advance"].append(a)
        s += a
    draft.kv_caching(False); target.kv_caching(False)
    stats["accepted_tokens"] = sum(stats["stage_accept_tokens"])
    stats["rejected_tokens"] = sum(stats["stage_tokens"]) - stats["accepted_tokens"]
    return f_hat, final_idx, stats


# --------------------------------------------------------------------------------------
# per-image ragged schedule + the reference's gamma controller (PARITY UNPINNED like the loop above)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def sd_generate_ragged(draft: RefVAR, target: RefVAR, vq: RefVQ, B: int, label_B: torch.Tensor, noise, cfg=1.5, gamma=2,
                       top_k=0, top_p=0.0, gamma_policy: str = "fixed"):
    """Spec of ``schedule='ragged'``: every image keeps its own stage pointer (and, under ``gamma_policy='reference'``, its own
    window length).  A round visits the groups of images that share (stage, gamma) in ascending order; each group runs the
    round of ``sd_generate`` on its sub-batch -- draft g stages, one block-causal target pass, per-token verify -- and every
    image of the group advances by ITS OWN accepted prefix: a_b = min(#leading window stages image b accepted whole + 1, g)
    (replaces the batch-global accept_length of models/var.py:1349-1350).  gamma controller (models/var.py:1352-1358): an
    image whose first drafted stage was not accepted whole shrinks its window by one, never below 1.

    The oracle keeps no ragged KV cache: the caches of a group are rebuilt by teacher-forcing its committed stages (the
    engine instead keeps per-image cache slots and runs the group through a slot map; the arithmetic is the same).
    Noise: per group, per stage, in the order the groups are visited: 'draft' (n*l,V) per drafted stage, 'u' (n*l) and
    'resample' (n*l,V) per verified stage.  Returns (f_hat, final idx per stage, stats)."""
    K, V = len(draft.patch_nums), draft.V
    HW = draft.patch_nums[-1]
    f_hat = torch.zeros(B, draft.Cvae, HW, HW)
    final = [torch.zeros(B, l, dtype=torch.int64) for l in draft.ls]
    stage, gammas = [0] * B, [gamma] * B
    stats = dict(rounds=0, target_passes=0, draft_stages=0, advance=[], image_rounds=[0] * B)

    def prefill(model, cond, ids, s):
        """rebuild the KV cache of the committed stages [0, s) of images ``ids``; returns (f_hat of the group, stage-s input map)"""
        model.kv_caching(False); model.kv_caching(True)
        fh = torch.zeros(len(ids), draft.Cvae, HW, HW)
        nm = None
        for si in range(s):
            x = model.first_map(cond) if si == 0 else model.embed_map(si, nm)
            model.blocks(x, cond, None)
            fh, nm = vq.next_input(si, fh, final[si][ids])
        return fh, nm

    while min(stage) < K:
        keys = sorted({(stage[b], gammas[b]) for b in range(B) if stage[b] < K})
        adv_round = []
        for s, gm in keys:
            ids = [b for b in range(B) if stage[b] == s and gammas[b] == gm]
            n, g = len(ids), min(gm, K - s)
            lab = label_B[ids]
            cond_d, cond_t = draft.cond(lab), target.cond(lab)
            fh0, nm0 = prefill(draft, cond_d, ids, s)
            prefill(target, cond_t, ids, s)
            fh = fh0.clone()
            maps, snaps, idx_d, mixed_d = [nm0], [], [], []
            for j in range(g):
                si = s + j
                x = draft.first_map(cond_d) if si == 0 else draft.embed_map(si, maps[j])
                mixed = cfg_mix(draft.get_logits(draft.blocks(x, cond_d, None), cond_d), n, cfg * si / (K - 1))
                idx = sample_with_noise_(mixed, noise.exponential("draft", n * draft.ls[si], V), top_k, top_p)
                fh, nm = vq.next_input(si, fh, idx)
                idx_d.append(idx); mixed_d.append(mixed); snaps.append(fh.clone()); maps.append(nm)
                stats["draft_stages"] += 1
            xs = [target.first_map(cond_t) if s + j == 0 else target.embed_map(s + j, maps[j]) for j in range(g)]
            logits_t = target.forward_window(s, xs, cond_t)
            stats["target_passes"] += 1
            out_idx, ok = [], torch.ones(n, g, dtype=torch.bool)
            for j in range(g):
                si, l = s + j, draft.ls[s + j]
                mixed_t = filter_top_k_top_p_(cfg_mix(logits_t[j], n, cfg * si / (K - 1)), top_k, top_p)
                u = noise.uniform("u", n * l)
                n_r = noise.exponential("resample", n * l, V)
                o, acc, _, _ = verify_tokens(mixed_t.view(-1, V), mixed_d[j].view(-1, V), idx_d[j].view(-1), u, n_r)
                out_idx.append(o.view(n, l))
                ok[:, j] = acc.view(n, l).all(dim=1)
            n_ok = ok.long().cumprod(dim=1).sum(dim=1).tolist()      # leading stages accepted whole, per image
            for i, b in enumerate(ids):
                a = min(n_ok[i] + 1, g)
                for j in range(a - 1):
                    final[s + j][b] = idx_d[j][i]
                final[s + a - 1][b] = out_idx[a - 1][i]
                base = snaps[a - 2][i:i + 1].clone() if a >= 2 else fh0[i:i + 1].clone()
                f_hat[b:b + 1], _ = vq.next_input(s + a - 1, base, out_idx[a - 1][i:i + 1])
                stage[b] = s + a
                stats["image_rounds"][b] += 1
                adv_round.append(a)
                if gamma_policy == "reference" and n_ok[i] == 0:
                    gammas[b] = max(1, gm - 1)
        stats["rounds"] += 1
        stats["advance"].append(sum(adv_round) / len(adv_round))
    draft.kv_caching(False); target.kv_caching(False)
    return f_hat, final, stats
