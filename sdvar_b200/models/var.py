"""VAR and SDVAR with the reference's Python API (reference models/var.py), executed by libsdvar_b200.

* ``VAR`` keeps the reference's constructor, parameter names/shapes (``var_d*.pth`` loads with
  ``strict=True``), ``autoregressive_infer_cfg`` (models/var.py:128-215), ``forward`` (:217-259) and
  ``get_logits`` (:119-125).  The modules below only HOLD parameters; all arithmetic is done by the CUDA
  engine (``sdvar_b200.engine.VarEngine``).  There is no CPU path: calling inference on a CPU model raises.
* ``SDVAR`` keeps ``sdvar_autoregressive_infer_cfg_parallel_v1`` (:1285-1383), ``..._sd_test3`` (:605-865) and
  the step methods (:871-1282), with the reference's defects D1-D11 resolved as specified in
  DESIGN.md "loop spec".  The reference's top-1/50% rule is kept as ``accept_rule='reference'``; the
  north-star rule ``u < min(1, p/q)`` + residual resample is ``accept_rule='speculative'`` (default).
"""
from __future__ import annotations

import os

import math
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from .. import _cabi
from ..engine import DeviceNoise, SingleGeneratorNoise, VarEngine
from .vqvae import VQVAE
from .quant import VectorQuantizer2


# ---------------------------------------------------------------------------------------------------
# parameter holders (names = checkpoint surface)
# ---------------------------------------------------------------------------------------------------
class _AttnParams(nn.Module):
    def __init__(self, C: int, H: int, l2: bool):
        super().__init__()
        if l2:
            self.scale_mul_1H11 = nn.Parameter(torch.full((1, H, 1, 1), 4.0).log())
        self.mat_qkv = nn.Linear(C, 3 * C, bias=False)
        self.q_bias, self.v_bias = nn.Parameter(torch.zeros(C)), nn.Parameter(torch.zeros(C))
        self.register_buffer("zero_k_bias", torch.zeros(C))
        self.proj = nn.Linear(C, C)


class _FfnParams(nn.Module):
    def __init__(self, C: int, hidden: int):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(C, hidden), nn.Linear(hidden, C)


class _BlockParams(nn.Module):
    def __init__(self, C: int, D: int, H: int, mlp_ratio: float, shared_aln: bool, l2: bool):
        super().__init__()
        self.attn = _AttnParams(C, H, l2)
        self.ffn = _FfnParams(C, round(C * mlp_ratio))
        self.shared_aln = shared_aln
        if shared_aln:
            self.ada_gss = nn.Parameter(torch.randn(1, 1, 6, C) / C ** 0.5)
        else:
            self.ada_lin = nn.Sequential(nn.SiLU(), nn.Linear(D, 6 * C))


class _HeadNorm(nn.Module):
    def __init__(self, C: int, D: int):
        super().__init__()
        self.ada_lin = nn.Sequential(nn.SiLU(), nn.Linear(D, 2 * C))


def _on_model_device(fn):
    """Run an inference entry with the model's GPU as the current CUDA device: libsdvar launches on the current device and
    stream, so a model living on cuda:1 must not be driven while cuda:0 is current."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        dev = self.device if isinstance(self, VAR) else self.target_model.device
        if dev.type != "cuda":
            raise _cabi.SdvarError("sdvar_b200 runs on sm_100a only: move the model to a CUDA device (no CPU fallback)")
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapped


def _trunc(t: torch.Tensor, std: float):
    nn.init.trunc_normal_(t, mean=0.0, std=std)


class VAR(nn.Module):
    def __init__(self, vae_local: VQVAE, num_classes=1000, depth=16, embed_dim=1024, num_heads=16, mlp_ratio=4.0,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, norm_eps=1e-6, shared_aln=False, cond_drop_rate=0.1,
                 attn_l2_norm=False, patch_nums=(1, 2, 3, 4, 5, 6, 8, 10, 13, 16), flash_if_available=True,
                 fused_if_available=True):
        super().__init__()
        assert embed_dim % num_heads == 0 and embed_dim // num_heads == 64, "the attention kernel is built for head_dim 64"
        assert mlp_ratio == 4.0
        self.Cvae, self.V = vae_local.Cvae, vae_local.vocab_size
        self.depth, self.C, self.D, self.num_heads = depth, embed_dim, embed_dim, num_heads
        self.cond_drop_rate, self.norm_eps, self.drop_path_rate = cond_drop_rate, norm_eps, drop_path_rate
        self.shared_aln, self.attn_l2_norm = shared_aln, attn_l2_norm
        self.prog_si = -1
        self.patch_nums: Tuple[int, ...] = tuple(patch_nums)
        self.ls = [pn * pn for pn in self.patch_nums]
        self.ends = [int(e) for e in np.cumsum(self.ls)]
        self.begins = [0] + self.ends[:-1]
        self.begin_ends = list(zip(self.begins, self.ends))
        self.L, self.first_l = self.ends[-1], self.ls[0]
        self.num_stages_minus_1 = len(self.patch_nums) - 1
        self.num_classes = num_classes
        self.rng: Optional[torch.Generator] = None   # created on the model's device on first use (var.py:50)
        # the VQVAE is deliberately kept out of .modules()/.state_dict(), like the reference's tuple proxy (var.py:54-55)
        self.vae_proxy: Tuple[VQVAE] = (vae_local,)
        self.vae_quant_proxy: Tuple[VectorQuantizer2] = (vae_local.quantize,)

        init_std = math.sqrt(1 / self.C / 3)
        self.word_embed = nn.Linear(self.Cvae, self.C)
        self.class_emb = nn.Embedding(num_classes + 1, self.C)
        self.pos_start = nn.Parameter(torch.empty(1, self.first_l, self.C))
        self.pos_1LC = nn.Parameter(torch.empty(1, self.L, self.C))
        self.lvl_embed = nn.Embedding(len(self.patch_nums), self.C)
        for t in (self.class_emb.weight, self.pos_start, self.pos_1LC, self.lvl_embed.weight):
            _trunc(t.data, init_std)
        if shared_aln:
            self.shared_ada_lin = nn.Sequential(nn.SiLU(), nn.Linear(self.D, 6 * self.C))
        else:
            self.shared_ada_lin = nn.Identity()
        self.blocks = nn.ModuleList([_BlockParams(self.C, self.D, num_heads, mlp_ratio, shared_aln, attn_l2_norm) for _ in range(depth)])
        lvl = torch.cat([torch.full((l,), i, dtype=torch.int64) for i, l in enumerate(self.ls)]).view(1, self.L)
        self.register_buffer("lvl_1L", lvl)
        d = lvl.view(1, self.L, 1)
        # kept only because it is part of the checkpoint surface; the kernels take the stage table instead
        self.register_buffer("attn_bias_for_masking", torch.where(d >= d.transpose(1, 2), 0.0, -torch.inf).reshape(1, 1, self.L, self.L).contiguous())
        self.head_nm = _HeadNorm(self.C, self.D)
        self.head = nn.Linear(self.C, self.V)
        self._engine = VarEngine(self)

    # ------------------------------------------------------------------ init (models/var.py:261-311)
    def init_weights(self, init_adaln=0.5, init_adaln_gamma=1e-5, init_head=0.02, init_std=0.02, conv_std_or_gain=0.02):
        if init_std < 0:
            init_std = (1 / self.C / 3) ** 0.5
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                _trunc(mod.weight.data, init_std)
                if mod.bias is not None:
                    mod.bias.data.zero_()
            elif isinstance(mod, nn.Embedding):
                _trunc(mod.weight.data, init_std)
        if init_head >= 0:
            self.head.weight.data.mul_(init_head)
            self.head.bias.data.zero_()
        self.head_nm.ada_lin[-1].weight.data.mul_(init_adaln)
        self.head_nm.ada_lin[-1].bias.data.zero_()
        for blk in self.blocks:
            blk.attn.proj.weight.data.div_(math.sqrt(2 * self.depth))
            blk.ffn.fc2.weight.data.div_(math.sqrt(2 * self.depth))
            if self.shared_aln:
                blk.ada_gss.data[:, :, 2:].mul_(init_adaln)
                blk.ada_gss.data[:, :, :2].mul_(init_adaln_gamma)
            else:
                blk.ada_lin[-1].weight.data[2 * self.C:].mul_(init_adaln)
                blk.ada_lin[-1].weight.data[:2 * self.C].mul_(init_adaln_gamma)
                blk.ada_lin[-1].bias.data.zero_()
        self.repack()      # .data writes do not bump tensor versions: drop the packed bf16 copy explicitly

    def repack(self):
        """Invalidate the engine's packed bf16 weights; call after mutating parameters through ``.data`` (EMA updates,
        manual surgery).  ``load_state_dict`` and ``init_weights`` do it themselves."""
        self._engine._packed_key = None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        self.repack()
        return super().load_state_dict(state_dict, strict=strict, assign=assign)

    def extra_repr(self) -> str:
        """reference models/var.py:313-314"""
        return f"drop_path_rate={self.drop_path_rate:g}"

    # ------------------------------------------------------------------ helpers
    @property
    def device(self):
        return self.pos_1LC.device

    def _rng(self, g_seed: Optional[int]) -> Optional[torch.Generator]:
        if g_seed is None:
            return None
        if self.rng is None or self.rng.device != self.device:
            self.rng = torch.Generator(device=self.device)
        self.rng.manual_seed(g_seed)
        return self.rng

    def _labels(self, B: int, label_B, rng) -> torch.Tensor:
        """models/var.py:147-150"""
        if label_B is None:
            prob = torch.full((1, self.num_classes), 1.0 / self.num_classes, device=self.device)
            return torch.multinomial(prob, num_samples=B, replacement=True, generator=rng).reshape(B)
        if isinstance(label_B, int):
            return torch.full((B,), self.num_classes if label_B < 0 else label_B, device=self.device, dtype=torch.int64)
        return label_B.to(self.device)

    def _cfg_scalars(self, cfg: float, stages: Sequence[int]):
        K1 = self.num_stages_minus_1
        t = [cfg * (si / K1) for si in stages]
        return [float(np.float32(1 + x)) for x in t], [float(np.float32(x)) for x in t]

    def _sample_stage(self, logits_2BLV, B, si, cfg, top_k, top_p, noise, in_ld=None, in_off=0, want_mixed=False):
        """K3 on one stage: returns (idx (B,l) int64 or None, mixed (B,l,V) or None)."""
        l, V = self.ls[si], self.V
        t1, t2 = self._cfg_scalars(cfg, [si])
        idx = torch.empty(B, l, dtype=torch.int64, device=self.device) if noise is not None else None
        mixed = torch.empty(B, l, V, dtype=torch.float32, device=self.device) if want_mixed else None
        thr = float(np.float32(1.0 - top_p)) if top_p > 0 else -1.0
        _cabi.sample_cfg_topk_topp(logits_2BLV, B, l, V, [0, l], t1, t2, top_k, thr, noise, idx, mixed, None, in_ld=in_ld, in_off=in_off)
        return idx, mixed

    def _stage_step(self, si, logits, B, cfg, top_k, top_p, n, more_smooth, rng, f_hat):
        """sample stage ``si`` from its logits and fold it into f_hat (models/var.py:199-211): (idx, f_hat, next_map)."""
        vq, K = self.vae_quant_proxy[0], len(self.patch_nums)
        if not more_smooth:
            idx, _ = self._sample_stage(logits, B, si, cfg, top_k, top_p, n)
            f_hat, next_map = vq.next_input_from_idx(si, f_hat, idx)
        else:  # visualisation-only branch (models/var.py:206-208), torch ops on the kernel's filtered logits
            idx, mixed = self._sample_stage(logits, B, si, cfg, top_k, top_p, n, want_mixed=True)
            ratio = si / self.num_stages_minus_1
            gum = -torch.empty_like(mixed).exponential_(generator=rng).log()
            y = ((mixed * (1 + ratio) + gum) / max(0.27 * (1 - ratio * 0.95), 0.005)).softmax(-1)
            h = (y @ vq.embedding.weight.unsqueeze(0)).transpose(1, 2).reshape(B, self.Cvae, self.patch_nums[si], self.patch_nums[si])
            f_hat, next_map = vq.get_next_autoregressive_input(si, K, f_hat, h.contiguous())
        return idx, f_hat, next_map

    def get_logits(self, h_BLC: torch.Tensor, cond_BD: torch.Tensor) -> torch.Tensor:
        """head(head_nm(h, cond)) in fp32 out (models/var.py:119-125): LN-modulate kernel + tcgen05 GEMM."""
        e = self._engine
        e.pack()
        n, L, C = h_BLC.shape
        silu = torch.empty(n, C, dtype=torch.bfloat16, device=self.device)
        _cabi.silu_bf16(cond_BD.float().contiguous(), silu)
        mod = torch.empty(n, 2 * C, device=self.device)
        E = _cabi.GemmEpilogue
        _cabi.gemm_bf16(silu, C, e.w_headnm, C, n, 2 * C, C, E(epilogue=_cabi.EPI_F32, bias=e.b_headnm.data_ptr(), out_f32=mod.data_ptr(), ldo=2 * C))
        xm = torch.empty(n * L, C, dtype=torch.bfloat16, device=self.device)
        _cabi.ln_modulate(h_BLC.float().contiguous().view(n * L, C), n * L, C, L, mod.data_ptr(), mod.data_ptr() + 4 * C, 2 * C, self.norm_eps, xm)
        out = torch.empty(n * L, self.V, device=self.device)
        _cabi.gemm_bf16(xm, C, e.w_head, C, n * L, self.V, C, E(epilogue=_cabi.EPI_F32, bias=e.b_head.data_ptr(), out_f32=out.data_ptr(), ldo=self.V))
        return out.view(n, L, self.V)

    # ------------------------------------------------------------------ baseline loop (models/var.py:128-215)
    @torch.no_grad()
    @_on_model_device
    def autoregressive_infer_cfg(self, B: int, label_B: Optional[Union[int, torch.LongTensor]], g_seed: Optional[int] = None,
                                 cfg=1.5, top_k=0, top_p=0.0, more_smooth=False, noise=None, return_tokens=False,
                                 record: Optional[dict] = None) -> torch.Tensor:
        """Returns the reconstructed image (B,3,H,W) in [0,1].  ``noise`` (optional) is a provider object with
        ``exponential(stream, rows, V)`` used by parity tests to inject pre-drawn Exp(1) noise; by default the
        model's generator is used exactly as torch.multinomial would use it."""
        rng = self._rng(g_seed)
        label_B = self._labels(B, label_B, rng)
        noise = noise or SingleGeneratorNoise(rng, self.device)
        vq, e = self.vae_quant_proxy[0], self._engine
        e.begin(B, label_B)
        K = len(self.patch_nums)
        f_hat = torch.zeros(B, self.Cvae, self.patch_nums[-1], self.patch_nums[-1], device=self.device)
        next_map, idxs = None, []
        for si in range(K):
            l = self.ls[si]
            if si == 0:
                e.put_first_map(l)
            else:
                e.put_embed_map(si, next_map, l)
            logits = e.forward([si])
            n = noise.exponential("target", B * l, self.V)
            if record is not None:
                record.setdefault("logits", []).append(logits.clone()); record.setdefault("noise", []).append(n)
            idx, f_hat, next_map = self._stage_step(si, logits, B, cfg, top_k, top_p, n, more_smooth, rng, f_hat)
            idxs.append(idx)
        img = self.vae_proxy[0].fhat_to_img(f_hat).add_(1).mul_(0.5)
        return (img, idxs, f_hat) if return_tokens else img

    # ------------------------------------------------------------------ teacher-forced (models/var.py:217-259)
    @torch.no_grad()
    @_on_model_device
    def forward(self, label_B: torch.LongTensor, x_BLCv_wo_first_l: torch.Tensor) -> torch.Tensor:
        """All-stage block-causal pass; returns logits (B, L, V).  Runs as one verify-style window over every stage.
        The reference drops labels at cond_drop_rate even in eval (var.py:226, not gated on training); training is
        out of scope here, so no label is dropped."""
        B = x_BLCv_wo_first_l.shape[0]
        e = self._engine
        e.begin(B, label_B, max_window_tokens=self.L)
        K = len(self.patch_nums)
        e.put_first_map(self.L, 0)
        for si in range(1, K):
            nm = x_BLCv_wo_first_l[:, self.begins[si] - self.first_l:self.ends[si] - self.first_l].float().transpose(1, 2).contiguous()
            e.put_embed_map(si, nm, self.L, self.begins[si])
        logits = e.forward(list(range(K)))
        return logits[:B].clone()


# ---------------------------------------------------------------------------------------------------
class SDVARInferenceState:
    """State of one speculative generation (reference: inner class at models/var.py:912-945).

    Committed state is per IMAGE (``stage[b]``, ``f_hat[b]``, ``final[s][b]``, the KV slots b / B+b of both engines); a round
    works on one GROUP of images that share a stage -- the whole batch under ``schedule='lockstep'``, a sub-batch under
    ``schedule='ragged'`` -- and ``current_stage`` / ``gamma`` / ``group`` describe that group."""

    def __init__(self, B, gamma, patch_nums, cfg, noise):
        self.B, self.gamma, self.patch_nums, self.cfg, self.noise = B, gamma, patch_nums, cfg, noise
        self.current_stage, self.total_stages = 0, len(patch_nums)
        self.accept_count = self.reject_count = self.target_calls = self.draft_stage_calls = self.rounds = 0
        self.top_k, self.top_p, self.more_smooth = 0, 0.0, False
        self.schedule, self.gamma_policy, self.record = "lockstep", "fixed", None
        self.verify_mode, self.lazy, self.lazy_skipped = "window", False, 0
        self.p_full = 0.0             # running estimate of P(a drafted stage is accepted whole), drives verify_mode='auto'
        self.f_hat = None             # committed f_hat (B,Cvae,HW,HW)
        self.final: List[torch.Tensor] = []     # committed tokens per stage, (B, l_s) int64
        self.stage = [0] * B          # per image: next stage to produce
        self.gammas = [gamma] * B     # per image window length (moves only under gamma_policy='reference')
        self.group: Optional[torch.Tensor] = None   # image ids of the current group (None = all, dense fast path)
        self.n = B                    # images in the current group
        # per-round scratch of the current group
        self.maps, self.snaps, self.out_idx = [], [], None
        self.win_xd = self.win_xt = self.win_d = None
        self.advance: List = []
        self.stage_tokens = [0] * len(patch_nums)
        self.stage_accept_tokens = [0] * len(patch_nums)

    # reference-compatible aliases
    @property
    def draft_f_hat(self):
        return self.f_hat

    @property
    def target_f_hat(self):
        return self.f_hat

    @property
    def final_idx(self) -> List[torch.Tensor]:
        return self.final

    def stats(self) -> dict:
        acc = sum(self.stage_accept_tokens)
        return dict(rounds=self.rounds, target_passes=self.target_calls, draft_stages=self.draft_stage_calls,
                    accepted_tokens=acc, rejected_tokens=sum(self.stage_tokens) - acc, advance=list(self.advance),
                    stage_accept_tokens=list(self.stage_accept_tokens), stage_tokens=list(self.stage_tokens),
                    schedule=self.schedule, gamma_policy=self.gamma_policy, verify_mode=self.verify_mode,
                    target_stages_skipped=self.lazy_skipped)


def _draw(noise, kind: str, stream: str, out: torch.Tensor):
    """fill the contiguous slice ``out`` from a noise provider; providers without ``out=`` support return a tensor to copy"""
    try:
        if kind == "exp":
            noise.exponential(stream, out.shape[0], out.shape[1], out=out)
        else:
            noise.uniform(stream, out.shape[0], out=out)
    except TypeError:
        out.copy_(noise.exponential(stream, out.shape[0], out.shape[1]) if kind == "exp" else noise.uniform(stream, out.shape[0]))


class SDVAR(nn.Module):
    LAZY_MIN_ROWS = 1024      # verify_mode='auto': rows (CFG included) from which a single-stage target pass is compute-bound on a B200
    OVERLAP_MAX_ROWS = int(os.environ.get("SDVAR_OVERLAP_ROWS", str(1 << 30)))   # lazy verification: draft and target work of a stage run concurrently up to this many rows (0 = never)

    def __init__(self, draft_model: VAR, target_model: VAR, similarity_thresh: float = 0.8):
        super().__init__()
        self.draft_model, self.target_model, self.similarity_thresh = draft_model, target_model, similarity_thresh
        assert draft_model.patch_nums == target_model.patch_nums, "draft and target must share the token pyramid"
        # The reference precomputes two dense LxL masks with O(L^2) python loops here (var.py:548-578, ~7 s) and
        # hard-codes the 256 px pyramid (D11).  The kernels take the stage table of whichever pyramid the models use.
        self.last_stats: Optional[dict] = None

    def _side_stream(self, dev) -> torch.cuda.Stream:
        if getattr(self, "_side", None) is None or self._side.device != torch.device(dev):
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    # models/var.py:580-601
    def init_param(self, model: VAR, B: int, label_B):
        e = model._engine
        e.pack()
        sos = cond_BD = e.class_emb[torch.cat((label_B, torch.full_like(label_B, model.num_classes)), dim=0)]
        lvl_pos = e.lvl_pos.unsqueeze(0)
        first_token_map = sos.unsqueeze(1).expand(2 * B, model.first_l, -1) + e.pos_start.unsqueeze(0) + lvl_pos[:, :model.first_l]
        first_f_hat = sos.new_zeros(B, model.Cvae, model.patch_nums[-1], model.patch_nums[-1])
        return sos, cond_BD, cond_BD, lvl_pos, first_token_map, first_f_hat

    # ------------------------------------------------------------------ hand-over variant (models/var.py:605-865)
    @torch.no_grad()
    @_on_model_device
    def sdvar_autoregressive_infer_cfg_sd_test3(self, B: int, label_B, g_seed: Optional[int] = None, cfg: float = 1.5,
                                                top_k: int = 0, top_p: float = 0.0, more_smooth: bool = False,
                                                entry_num: int = 10, sd_mask: int = 0, noise=None, return_tokens=False):
        """Draft runs stages [0,entry_num), the target runs [entry_num,K) from the draft's f_hat.  With sd_mask=0 the
        target starts with an EMPTY KV cache and never sees the drafted prefix (var.py:817-825) -- reproduced as is.
        sd_mask 1..5 run the prefix through the blocks and then discard the result (var.py:809-811); that branch has
        no defined semantics and is refused."""
        if sd_mask != 0:
            raise NotImplementedError("sd_mask != 0 discards the masked pass in the reference (models/var.py:809-811); not supported")
        D, T = self.draft_model, self.target_model
        rng = T._rng(g_seed)
        label_B = T._labels(B, label_B, rng)
        noise = noise or SingleGeneratorNoise(rng, T.device)
        K = len(T.patch_nums)
        f_hat = torch.zeros(B, T.Cvae, T.patch_nums[-1], T.patch_nums[-1], device=T.device)
        next_map, idxs = None, []
        for model, lo, hi in ((D, 0, min(entry_num, K)), (T, min(entry_num, K), K)):
            if lo >= hi:
                continue
            e = model._engine
            e.begin(B, label_B)
            for si in range(lo, hi):
                l = model.ls[si]
                e.put_first_map(l) if si == 0 else e.put_embed_map(si, next_map, l)
                logits = e.forward([si], check_position=False)
                idx, f_hat, next_map = model._stage_step(si, logits, B, cfg, top_k, top_p, noise.exponential("target", B * l, model.V),
                                                         more_smooth, rng, f_hat)
                idxs.append(idx)
        img = T.vae_proxy[0].fhat_to_img(f_hat).add_(1).mul_(0.5)
        return (img, idxs, f_hat) if return_tokens else img

    # ------------------------------------------------------------------ draft -> verify loop (models/var.py:871-1383)
    def _initialize_inference_state(self, B: int, label_B, g_seed: Optional[int], cfg: float, gamma: int, noise=None):
        D, T = self.draft_model, self.target_model
        rng = T._rng(g_seed)
        label_B = T._labels(B, label_B, rng)
        state = SDVARInferenceState(B, gamma, T.patch_nums, cfg, noise or DeviceNoise(g_seed, T.device))
        state.label_B = label_B
        K = len(T.patch_nums)
        win = max(sum(T.ls[s:s + gamma]) for s in range(K))
        D._engine.begin(B, label_B)
        T._engine.begin(B, label_B, max_window_tokens=win)
        dev = T.device
        state.f_hat = torch.zeros(B, T.Cvae, T.patch_nums[-1], T.patch_nums[-1], device=dev)
        state.final = [torch.zeros(B, l, dtype=torch.int64, device=dev) for l in T.ls]
        state.ws = torch.zeros(_cabi.verify_workspace_ints(B, min(gamma, K)), dtype=torch.int32, device=dev)
        return state

    def _window(self, state: SDVARInferenceState):
        """(stages, per-stage token offsets, window length) of the current group's round"""
        T = self.target_model
        g = min(state.gamma, state.total_stages - state.current_stage)
        stages = list(range(state.current_stage, state.current_stage + g))
        offs = [0]
        for si in stages:
            offs.append(offs[-1] + T.ls[si])
        return stages, offs, offs[-1]

    def _group_f_hat(self, state: SDVARInferenceState) -> torch.Tensor:
        return state.f_hat.clone() if state.group is None else state.f_hat.index_select(0, state.group)

    def draft_generate_batch(self, state: SDVARInferenceState, B: int, upto: Optional[int] = None) -> List[torch.Tensor]:
        """Draft g = min(gamma, K - stage) stages incrementally for the current group (models/var.py:949-1024).  Each stage's
        input is area_down(f_hat) of the previous DRAFTED stage (fixes D2), stage 0 is the sos map (D3); K3 writes the draft's
        tokens and its mixed+filtered logits straight into the window-shaped buffers the verify kernel reads; f_hat
        snapshots are kept for the commit.  ``upto`` (lazy verification) drafts only the first ``upto`` stages of the window
        now; ``_draft_stage`` continues on demand."""
        D = self.draft_model
        e = D._engine
        stages, offs, Lw = self._window(state)
        n, V, dev = state.n, D.V, D.device
        state.snaps, state.draft_tokens = [], []
        if not stages:
            return state.draft_tokens
        state.draft_smap = None if state.group is None else e.slot_map(state.group)
        if state.lazy:     # stage-by-stage verification reads each stage on its own: per-stage dense buffers
            state.st_xd = [torch.empty(n, D.ls[si], V, dtype=torch.float32, device=dev) for si in stages]
            state.st_d = [torch.empty(n, D.ls[si], dtype=torch.int64, device=dev) for si in stages]
        else:
            state.win_xd = torch.empty(n, Lw, V, dtype=torch.float32, device=dev)
            state.win_d = torch.empty(n, Lw, dtype=torch.int64, device=dev)
        state.draft_fh = self._group_f_hat(state)
        nm = None
        if stages[0] > 0:      # stage input rebuilt from the committed f_hat (same kernel as the K5 step => same bits)
            pn = D.patch_nums[stages[0]]
            nm = torch.empty(n, D.Cvae, pn, pn, dtype=torch.float32, device=dev)
            _cabi.vq_area_down(state.draft_fh, n, D.patch_nums[-1], pn, D.Cvae, nm)
        state.maps = [nm]
        for j in range(len(stages) if upto is None else min(upto, len(stages))):
            self._draft_stage(state, j)
        return state.draft_tokens

    def _draft_pass(self, state: SDVARInferenceState, j: int) -> torch.Tensor:
        """transformer pass of draft stage j of the current window (no allocation: engine-owned buffers only, so it may run on a
        side stream); returns the raw logits (2n, l, V)"""
        D = self.draft_model
        e = D._engine
        stages, offs, Lw = self._window(state)
        si = stages[j]
        l = D.ls[si]
        assert j == len(state.draft_tokens)
        smap = state.draft_smap
        e.put_first_map(l, slot_map=smap) if si == 0 else e.put_embed_map(si, state.maps[j], l)
        return e.forward([si], slot_map=smap)

    def _draft_finish(self, state: SDVARInferenceState, j: int, logits: torch.Tensor):
        """noise draw, K3 (tokens + the draft's mixed / filtered logits), K5 (f_hat, next stage input) of draft stage j"""
        D = self.draft_model
        vq = D.vae_quant_proxy[0]
        stages, offs, Lw = self._window(state)
        si, n, V = stages[j], state.n, D.V
        l = D.ls[si]
        thr = float(np.float32(1.0 - state.top_p)) if state.top_p > 0 else -1.0
        nz = state.noise.exponential("draft", n * l, V)
        t1, t2 = D._cfg_scalars(state.cfg, [si])
        if state.lazy:
            _cabi.sample_cfg_topk_topp(logits, n, l, V, [0, l], t1, t2, state.top_k, thr, nz, state.st_d[j], state.st_xd[j], None)
            idx = state.st_d[j]
        else:
            _cabi.sample_cfg_topk_topp(logits, n, l, V, [0, l], t1, t2, state.top_k, thr, nz, state.win_d, state.win_xd, None,
                                       out_ld=Lw, out_off=offs[j])
            idx = state.win_d[:, offs[j]:offs[j + 1]].contiguous()
        state.draft_fh, nm = vq.next_input_from_idx(si, state.draft_fh, idx)
        state.draft_tokens.append(idx); state.snaps.append(state.draft_fh.clone())
        state.maps.append(nm if si != state.total_stages - 1 else None)
        state.draft_stage_calls += 1

    def _draft_stage(self, state: SDVARInferenceState, j: int, skip: bool = False):
        """draft stage j of the current window (must follow stage j-1).  ``skip``: only make the stage's noise draw, so the
        'draft' stream stays where the loop spec puts it when lazy verification never needs this stage."""
        if skip:
            D = self.draft_model
            stages, _, _ = self._window(state)
            state.noise.exponential("draft", state.n * D.ls[stages[j]], D.V)
            return
        self._draft_finish(state, j, self._draft_pass(state, j))

    def target_verify_batch(self, draft_tokens: List[torch.Tensor], state: SDVARInferenceState, B: int):
        """ONE block-causal target pass over the g drafted stages on top of the KV cache of accepted stages
        (models/var.py:1026-1158 intent; fixes D1,D5,D6), then ONE K3 launch that CFG-mixes and filters the whole window with
        its per-stage strengths: returns ([mixed+filtered target logits (n,l,V) per stage, views of one window buffer], g)."""
        if not draft_tokens:
            return [], 0
        T = self.target_model
        e = T._engine
        stages, offs, Lw = self._window(state)
        g, n = len(stages), state.n
        smap = None if state.group is None else e.slot_map(state.group)
        for j, si in enumerate(stages):
            e.put_first_map(Lw, offs[j], slot_map=smap) if si == 0 else e.put_embed_map(si, state.maps[j], Lw, offs[j])
        logits = e.forward(stages, slot_map=smap)
        state.target_calls += 1
        t1, t2 = T._cfg_scalars(state.cfg, stages)
        thr = float(np.float32(1.0 - state.top_p)) if state.top_p > 0 else -1.0
        state.win_xt = torch.empty(n, Lw, T.V, dtype=torch.float32, device=T.device)
        _cabi.sample_cfg_topk_topp(logits, n, Lw, T.V, offs, t1, t2, state.top_k, thr, None, None, state.win_xt, None)
        return [state.win_xt[:, offs[j]:offs[j + 1]] for j in range(g)], g

    def lazy_verify_batch(self, draft_tokens: List[torch.Tensor], state: SDVARInferenceState, B: int) -> int:
        """``verify_mode='lazy'``: the window is drafted and verified STAGE BY STAGE with early exit -- one single-stage target
        pass, K3 filter and K4 launch per stage; the next stage is only drafted and run through the target if every token of
        this one was accepted (its noise is drawn either way).  A window pass equals the incremental passes bit for bit (tests/test_engine_gpu.py::
        test_window_pass_equals_incremental), every noise draw of the window is still made in the loop spec's order, and a stage
        behind the first rejected one is never committed anyway: tokens, accept flags, advances and images are IDENTICAL to the
        one-pass window verification; only target work that the window pass would have thrown away is not done.  When target
        passes are compute-bound (large batch) a rejected window then costs one stage, not g.  Lock-step schedule only."""
        T = self.target_model
        e = T._engine
        stages, offs, Lw = self._window(state)
        g, n, V, dev = len(stages), state.n, T.V, T.device
        assert state.group is None and state.schedule == "lockstep"
        u = torch.empty(n * Lw, dtype=torch.float32, device=dev)
        nr = torch.empty(n * Lw, V, dtype=torch.float32, device=dev)
        for j in range(g):       # all draws of the window, in the loop spec's order, whether or not the stage gets verified
            _draw(state.noise, "uni", "u", u[n * offs[j]:n * offs[j + 1]])
            _draw(state.noise, "exp", "resample", nr[n * offs[j]:n * offs[j + 1]])
        thr = float(np.float32(1.0 - state.top_p)) if state.top_p > 0 else -1.0
        out = torch.empty(n, Lw, dtype=torch.int64, device=dev)
        na_all = torch.zeros(n, g, dtype=torch.int32)
        n_ok = 0
        for j, si in enumerate(stages):
            l = T.ls[si]
            need_draft = j >= len(state.draft_tokens)         # lazy drafting: stage j is drafted only now that stage j-1 survived
            side = None
            if need_draft and 2 * n * l <= self.OVERLAP_MAX_ROWS:
                # The draft's and the target's work on stage j read the same committed state and write disjoint buffers, so they
                # run CONCURRENTLY: the draft (pass, noise, K3, K5) on a side stream, the target pass and its K3 on the main one;
                # they meet again at the verify kernel.  Every side-stream episode starts with side.wait_stream(main) and ends
                # with main.wait_stream(side), which also orders the caching allocator's reuse of blocks across the two
                # streams.  Same kernels on the same data: identical results.  (B=64: +2 %, B=8: +12 % images/s.)
                main = torch.cuda.current_stream()
                side = self._side_stream(dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    self._draft_stage(state, j)
            elif need_draft:
                self._draft_stage(state, j)
            e.put_first_map(l) if si == 0 else e.put_embed_map(si, state.maps[j], l)
            logits = e.forward([si])
            state.target_calls += 1
            t1, t2 = T._cfg_scalars(state.cfg, [si])
            xt = torch.empty(n, l, V, dtype=torch.float32, device=dev)
            _cabi.sample_cfg_topk_topp(logits, n, l, V, [0, l], t1, t2, state.top_k, thr, None, None, xt, None)
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)
            o = torch.empty(n, l, dtype=torch.int64, device=dev)
            acc = torch.empty(n, l, dtype=torch.uint8, device=dev)
            fr = torch.empty(n, 1, dtype=torch.int32, device=dev); na = torch.empty(n, 1, dtype=torch.int32, device=dev)
            st = torch.empty(n, dtype=torch.int32, device=dev); summ = torch.empty(4, dtype=torch.int32, device=dev)
            _cabi.verify_accept_resample(xt, state.st_xd[j], state.st_d[j], u[n * offs[j]:n * offs[j + 1]], nr[n * offs[j]:n * offs[j + 1]],
                                         n, l, V, [0, l], o, acc, None, None, fr, na, st, summ, state.ws)
            out[:, offs[j]:offs[j + 1]] = o
            h = torch.cat((summ, na.flatten())).cpu()         # one host sync per verified stage
            na_all[:, j] = h[4:]
            if int(h[0]) < 1:                                 # some image rejected a token of this stage: it is repaired, the rest is moot
                break
            n_ok += 1
        for j in range(len(state.draft_tokens), g):          # never needed: keep the 'draft' noise stream in step with the spec
            self._draft_stage(state, j, skip=True)
        state.lazy_skipped += g - min(n_ok + 1, g)
        state.out_idx = out
        state.last_n_ok = torch.full((n,), n_ok, dtype=torch.int32)
        return self._advance_from_counts(state, stages, state.last_n_ok, na_all, n_ok)

    def _advance_from_counts(self, state, stages, n_ok_per_image, na_host, lockstep_n_ok):
        """statistics + advance lengths from the per-(image, stage) accept counts of the round"""
        T = self.target_model
        g, n = len(stages), state.n
        if state.schedule == "lockstep":
            for j, si in enumerate(stages):   # stages after the first failing one were conditioned on unrepaired tokens: not counted
                if j > lockstep_n_ok:
                    break
                state.stage_tokens[si] += n * T.ls[si]
                state.stage_accept_tokens[si] += int(na_host[:, j].sum())
            return min(lockstep_n_ok + 1, g)
        adv = []
        for i in range(n):
            ok = int(n_ok_per_image[i])
            for j, si in enumerate(stages):
                if j > ok:
                    break
                state.stage_tokens[si] += T.ls[si]
                state.stage_accept_tokens[si] += int(na_host[i, j])
            adv.append(min(ok + 1, g))
        return adv

    def speculative_token_matching(self, draft_tokens, target_logits, state: SDVARInferenceState, B: int):
        """K4 over the whole window in ONE launch: accept u*q[d] < p[d], residual resample on reject, per-(image, stage)
        first-reject scan.  Returns the number of stages to commit: lock-step a = min(#leading stages with no reject in ANY
        image + 1, g) (an int, the reference's ``accept_length``); under ``schedule='ragged'`` one such number PER IMAGE
        (a list).  The last committed stage keeps the accepted draft tokens plus the target's repairs."""
        T = self.target_model
        stages, offs, Lw = self._window(state)
        g, n, V, dev = len(stages), state.n, T.V, T.device
        u = torch.empty(n * Lw, dtype=torch.float32, device=dev)
        nr = torch.empty(n * Lw, V, dtype=torch.float32, device=dev)
        for j in range(g):       # the per-stage draws of the loop spec, laid end to end ("stage-major")
            _draw(state.noise, "uni", "u", u[n * offs[j]:n * offs[j + 1]])
            _draw(state.noise, "exp", "resample", nr[n * offs[j]:n * offs[j + 1]])
        out = torch.empty(n, Lw, dtype=torch.int64, device=dev)
        acc = torch.empty(n, Lw, dtype=torch.uint8, device=dev)
        fr = torch.empty(n, g, dtype=torch.int32, device=dev); na = torch.empty(n, g, dtype=torch.int32, device=dev)
        st = torch.empty(n, dtype=torch.int32, device=dev); summ = torch.empty(4, dtype=torch.int32, device=dev)
        _cabi.verify_accept_resample(state.win_xt, state.win_xd, state.win_d, u, nr, n, Lw, V, offs, out, acc, None, None, fr, na,
                                     st, summ, state.ws, stage_major_aux=True)
        state.out_idx = out
        if state.record is not None:
            state.record.setdefault("rounds", []).append(dict(
                stage=stages[0], seg=list(offs), group=None if state.group is None else state.group.clone(), xt=state.win_xt.clone(),
                xd=state.win_xd.clone(), d=state.win_d.clone(), u=u.clone(), noise=nr.clone(), out=out.clone(), accept=acc.clone(),
                first_reject=fr.clone(), n_accept=na.clone(), accepted_stages=st.clone(), summary=summ.clone()))
        h = torch.cat((summ, st, na.flatten())).cpu()      # the one host sync of the round
        st_h, na_h = h[4:4 + n], h[4 + n:].view(n, g)
        res = self._advance_from_counts(state, stages, st_h, na_h, int(h[0]))
        state.last_n_ok = st_h
        return res

    def basic_token_matching(self, draft_tokens, target_logits, state: SDVARInferenceState, B: int) -> int:
        """Reference rule (models/var.py:1160-1227): stage accepted iff the batch-mean top-1 match rate >= 0.5, stop at
        the first rejected stage.  Departure (D8): the rejected stage is repaired by sampling it from the target's
        distribution instead of breaking out of the loop with a truncated image.  The rule is batch-global by definition
        (D9), so it only runs under the lock-step schedule."""
        T = self.target_model
        stages, offs, Lw = self._window(state)
        g, n, V, dev = len(stages), state.n, T.V, T.device
        assert state.schedule == "lockstep", "accept_rule='reference' couples the batch (models/var.py:1203): lock-step only"
        match = torch.empty(n, Lw, dtype=torch.uint8, device=dev)
        nm = torch.empty(n, g, dtype=torch.int32, device=dev)
        _cabi.verify_top1(state.win_xt, state.win_d, n, Lw, V, offs, match, nm)
        counts = nm.cpu().sum(dim=0).tolist()
        n_ok = 0
        for j, si in enumerate(stages):
            l = T.ls[si]
            if j == n_ok:
                state.stage_tokens[si] += n * l
                if counts[j] / float(n * l) >= 0.5:
                    n_ok += 1
                    state.stage_accept_tokens[si] += counts[j]
        out = state.win_d.clone()
        for j, si in enumerate(stages):
            l = T.ls[si]
            u = state.noise.uniform("u", n * l)                      # drawn to keep the stream positions rule-independent
            nr = state.noise.exponential("resample", n * l, V)
            if j == n_ok:
                # the window buffer is already mixed+filtered: sample it with K3 as x = x*1 - 0*0
                x = state.win_xt[:, offs[j]:offs[j + 1]].contiguous()
                both = torch.cat((x, torch.zeros_like(x)), 0)
                _cabi.sample_cfg_topk_topp(both, n, l, V, [0, l], [1.0], [0.0], 0, -1.0, nr, out, None, None, out_ld=Lw, out_off=offs[j])
        state.out_idx = out
        state.last_n_ok = torch.full((n,), n_ok, dtype=torch.int32)
        return min(n_ok + 1, g)

    def advanced_token_matching(self, draft_tokens, target_logits, state: SDVARInferenceState, B: int) -> int:
        """The reference's placeholder (models/var.py:1229-1243) returns the basic rule's result; so does this."""
        return self.basic_token_matching(draft_tokens, target_logits, state, B)

    def update_state_with_accepted_tokens(self, draft_tokens, accept_length, state: SDVARInferenceState, B: int):
        """Commit ``accept_length`` stages of the current group (models/var.py:1245-1282): f_hat = snapshot after the last
        unmodified stage + the final tokens of the last committed stage (fixes the double add D7); both KV caches roll back
        to the end of the last committed stage (D4).  ``accept_length`` is an int (whole group) or a per-image list
        (ragged): images are then committed in sub-groups of equal length."""
        T, D = self.target_model, self.draft_model
        vq = T.vae_quant_proxy[0]
        stages, offs, Lw = self._window(state)
        s, n = state.current_stage, state.n
        ids = list(range(state.B)) if state.group is None else state.group.tolist()
        if isinstance(accept_length, int):
            parts = [(accept_length, None)] if accept_length > 0 else []
        else:
            parts = [(a, [i for i in range(n) if accept_length[i] == a]) for a in sorted(set(accept_length)) if a > 0]
        for a, rows in parts:
            sel = None if rows is None else torch.tensor(rows, device=T.device, dtype=torch.int64)
            dst = state.group if rows is None else torch.tensor([ids[i] for i in rows], device=T.device, dtype=torch.int64)
            pick = (lambda t: t) if sel is None else (lambda t: t.index_select(0, sel))
            last = pick(state.out_idx[:, offs[a - 1]:offs[a]]).contiguous()
            base = pick(state.snaps[a - 2]).clone() if a >= 2 else (self._group_f_hat(state) if sel is None else
                                                                     state.f_hat.index_select(0, dst))
            fh, _ = vq.next_input_from_idx(s + a - 1, base, last)
            if dst is None:
                for j in range(a - 1):
                    state.final[s + j].copy_(draft_tokens[j])
                state.final[s + a - 1].copy_(last)
                state.f_hat = fh
            else:
                for j in range(a - 1):
                    state.final[s + j].index_copy_(0, dst, pick(draft_tokens[j]))
                state.final[s + a - 1].index_copy_(0, dst, last)
                state.f_hat.index_copy_(0, dst, fh)
            for i in (range(n) if rows is None else rows):
                state.stage[ids[i]] = s + a
        if state.group is None and isinstance(accept_length, int) and accept_length > 0:
            D._engine.kv_truncate(D.ends[s + accept_length - 1])     # sub-batch passes address the cache by stage position instead
            T._engine.kv_truncate(T.ends[s + accept_length - 1])

    @torch.no_grad()
    @_on_model_device
    def sdvar_autoregressive_infer_cfg_parallel_v1(self, B: int, label_B=None, g_seed: Optional[int] = None, cfg: float = 1.5,
                                                   gamma: int = 2, top_k: int = 0, top_p: float = 0.0, more_smooth: bool = False,
                                                   accept_rule: str = "speculative", schedule: str = "lockstep",
                                                   gamma_policy: str = "fixed", noise=None, return_tokens: bool = False,
                                                   record: Optional[dict] = None, _bound: Optional[str] = None,
                                                   verify_mode: str = "window"):
        """while stage < K: draft g stages -> one target pass -> verify -> commit a prefix (models/var.py:1285-1383).
        Returns the image (B,3,H,W) in [0,1]; acceptance statistics are left in ``self.last_stats``.

        schedule      'lockstep' (default; the reference's batch-global ``accept_length``, var.py:1349-1350): the whole batch
                      advances by the minimum over images.  'ragged' (SURVEY.md 8f #2): every image keeps its own stage
                      pointer and advances by its own accepted prefix; each round runs one dense pass per group of images that
                      share a stage, addressed through a slot map, so accepted images do not wait for rejected ones.
        gamma_policy  'fixed' (default) or 'reference' = the reference's controller (var.py:1352-1372): after a round in which
                      no drafted stage survived intact the window shrinks by one (never below 1, never grows back).
        more_smooth   accepted and stored like the reference does (var.py:1315); the drafting / verification path never reads
                      it there either.
        verify_mode   how the target verifies a drafted window -- the RESULT is the same bit for bit, only the work differs:
                      'window' (default) one block-causal target pass over all g stages; 'lazy' stage by stage with early exit
                      (``lazy_verify_batch``); 'auto' lazy when the first stage alone already fills the GPU (2n*l_s >=
                      ``LAZY_MIN_ROWS`` rows: its pass is compute-bound, so a second stage costs its full FLOPs and is pure loss
                      when the first one gets repaired), or when fewer than half of the recently verified stages were accepted
                      whole (running estimate ``p_full``: in the latency-bound regime a window round costs three passes -- two
                      draft, one target -- and commits 1+p stages, a lazy round two passes per committed stage, so the window
                      pays only for p > 1/2); one window pass otherwise.  Lock-step + speculative rule only; anything else
                      verifies by window.
        record        optional dict: receives every round's verify inputs and outputs (loop-replay tests).
        _bound        measurement only (bench.py 'bounds'): 'accept_all' commits every drafted window whole, 'reject_all' commits
                      one stage per round, whatever the verify kernel said -- the two ends of the acceptance schedule."""
        assert accept_rule in ("speculative", "reference") and schedule in ("lockstep", "ragged") and gamma_policy in ("fixed", "reference")
        assert gamma >= 1 and _bound in (None, "accept_all", "reject_all") and verify_mode in ("window", "lazy", "auto")
        state = self._initialize_inference_state(B, label_B, g_seed, cfg, gamma, noise)
        can_lazy = verify_mode != "window" and schedule == "lockstep" and accept_rule == "speculative" and record is None and _bound is None
        state.verify_mode = verify_mode if can_lazy else "window"
        state.top_k, state.top_p, state.more_smooth = top_k, top_p, more_smooth
        state.schedule, state.gamma_policy, state.record = schedule, gamma_policy, record
        match = self.speculative_token_matching if accept_rule == "speculative" else self.basic_token_matching
        K, dev = state.total_stages, self.target_model.device
        while min(state.stage) < K:
            if schedule == "lockstep":
                groups = [(state.stage[0], state.gammas[0], None)]
            else:
                keys = sorted({(state.stage[b], state.gammas[b]) for b in range(B) if state.stage[b] < K})
                groups = [(s, gm, [b for b in range(B) if state.stage[b] == s and state.gammas[b] == gm]) for s, gm in keys]
            round_adv = []
            for s, gm, members in groups:
                state.current_stage, state.gamma = s, gm
                state.group = None if members is None or len(members) == B else torch.tensor(members, device=dev, dtype=torch.int64)
                state.n = B if state.group is None else len(members)
                ids = list(range(B)) if state.group is None else members
                state.lazy = can_lazy and min(gm, K - s) > 1 and (
                    verify_mode == "lazy" or 2 * state.n * self.target_model.ls[s] >= self.LAZY_MIN_ROWS or state.p_full <= 0.5)
                draft_tokens = self.draft_generate_batch(state, state.n, upto=0 if state.lazy else None)
                if state.lazy:
                    accept_length = self.lazy_verify_batch(draft_tokens, state, state.n)
                else:
                    target_logits, _ = self.target_verify_batch(draft_tokens, state, state.n)
                    accept_length = match(draft_tokens, target_logits, state, state.n)
                if _bound == "accept_all":
                    accept_length, state.out_idx = len(draft_tokens), state.win_d
                elif _bound == "reject_all":
                    accept_length = 1
                self.update_state_with_accepted_tokens(draft_tokens, accept_length, state, state.n)
                if can_lazy and _bound is None:     # every verified stage is one observation: the n_ok leading ones held, the next one did not
                    n_ok = int(min(state.last_n_ok.tolist()))
                    for smp in [1.0] * n_ok + ([0.0] if n_ok < min(gm, K - s) else []):
                        state.p_full = 0.5 * state.p_full + 0.5 * smp
                per_img = [accept_length] * state.n if isinstance(accept_length, int) else accept_length
                round_adv += per_img
                if gamma_policy == "reference":      # var.py:1352-1358: shrink the window after a round with no intact stage
                    n_ok = state.last_n_ok.tolist()
                    if schedule == "lockstep":
                        if min(n_ok) == 0:
                            state.gammas = [max(1, gm - 1)] * B
                    else:
                        for i, b in enumerate(ids):
                            if n_ok[i] == 0:
                                state.gammas[b] = max(1, gm - 1)
            state.accept_count += min(round_adv)
            state.rounds += 1
            state.advance.append(round_adv[0] if schedule == "lockstep" else round(sum(round_adv) / len(round_adv), 3))
        state.current_stage = K
        self.last_stats = state.stats()
        img = self.target_model.vae_proxy[0].fhat_to_img(state.f_hat).add_(1).mul_(0.5)
        return (img, state.final, state.f_hat) if return_tokens else img
