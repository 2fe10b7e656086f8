"""Device-side execution engine of one VAR model: packed bf16 weights, preallocated KV ring, workspaces,
and the per-pass launch through ``sdvar_var_forward`` (include/sdvar_b200.h).

PyTorch is used here for device memory, streams and a handful of gathers (class-embedding lookup); all
arithmetic on the path is done by libsdvar_b200.  There is no CPU fallback: constructing an engine on a
non-CUDA device raises.

Layout in HBM (per model, batch B, imgs = 2B for CFG):
  weights   bf16, row-major (out_features, in_features) exactly as nn.Linear stores them (K-major for UMMA)
  KV cache  per block: K (imgs, H, Lmax, 64) bf16, V^T (imgs, H, 64, Lmax_pad) bf16, zero-initialised; "append"
            is the QKV epilogue writing at kv_len, rollback is ``kv_len = n`` (fixes reference defect D4)
  x         (imgs*Lq_max, C) fp32 residual stream; xm/q/attn (imgs*Lq_max, C) bf16; hidden (.., 4C) bf16
  logits    (imgs*Lq_max, V) fp32
  ada       (imgs, depth*6C) fp32 adaLN table computed ONCE per generation (the reference recomputes it every
            stage, models/basic_var.py:156); head_mod (imgs, 2C) fp32
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import torch

from . import _cabi


class VarEngine:
    def __init__(self, var_module):
        self.m = var_module
        self._packed_key = None
        self._ws_key = None
        self.kv_len = 0
        # stage-level CUDA graphs: a dense pass over given stages at a given cache position always
        # launches the same kernels on the same buffers, so it is captured on its second use and replayed afterwards
        self.use_graphs = os.environ.get("SDVAR_GRAPHS", "1") != "0"
        self._graphs, self._graph_seen = {}, {}

    # ------------------------------------------------------------------ weights
    def _key(self):
        ps = list(self.m.parameters())
        return (str(ps[0].device), sum(p._version for p in ps), sum(p.data_ptr() & 0xFFFF for p in ps[:4]))

    def pack(self, force: bool = False):
        m = self.m
        key = self._key()
        if not force and key == self._packed_key:
            return
        dev = m.pos_1LC.device
        if dev.type != "cuda":
            raise _cabi.SdvarError("sdvar_b200 runs on sm_100a only: move the model to a CUDA device (no CPU fallback)")
        rc = _cabi.lib().sdvar_arch_check(dev.index or 0)
        if rc != 0:
            raise _cabi.SdvarError(_cabi.lib().sdvar_last_error().decode())
        self.dev = dev
        C, H, D = m.C, m.num_heads, m.depth
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        self.keep = []  # owns every packed tensor
        w = _cabi.VarWeights()
        w.depth, w.C, w.H, w.V, w.Cvae, w.l2norm = D, C, H, m.V, m.Cvae, int(m.attn_l2_norm)
        w.eps = m.norm_eps
        w.attn_scale = 1.0 if m.attn_l2_norm else 0.25 / math.sqrt(C // H)
        ada_w, ada_b = [], []
        for i, blk in enumerate(m.blocks):
            a = blk.attn
            t = dict(w_qkv=bf(a.mat_qkv.weight), b_qkv=f32(torch.cat((a.q_bias, a.zero_k_bias, a.v_bias))),
                     scale_mul=f32(a.scale_mul_1H11.view(-1)) if m.attn_l2_norm else None,
                     w_proj=bf(a.proj.weight), b_proj=f32(a.proj.bias), w_fc1=bf(blk.ffn.fc1.weight), b_fc1=f32(blk.ffn.fc1.bias),
                     w_fc2=bf(blk.ffn.fc2.weight), b_fc2=f32(blk.ffn.fc2.bias))
            self.keep.append(t)
            for k, v in t.items():
                getattr(w, k)[i] = 0 if v is None else v.data_ptr()
            if not m.shared_aln:
                ada_w.append(blk.ada_lin[1].weight); ada_b.append(blk.ada_lin[1].bias)
        if m.shared_aln:
            self.w_ada = bf(m.shared_ada_lin[1].weight); self.b_ada = f32(m.shared_ada_lin[1].bias)
            self.ada_gss = f32(torch.stack([b.ada_gss.view(6 * C) for b in m.blocks]))  # (depth, 6C)
        else:
            self.w_ada = bf(torch.cat(ada_w, 0)); self.b_ada = f32(torch.cat(ada_b, 0))     # (depth*6C, C)
        self.w_headnm = bf(m.head_nm.ada_lin[1].weight); self.b_headnm = f32(m.head_nm.ada_lin[1].bias)
        self.w_head = bf(m.head.weight); self.b_head = f32(m.head.bias)
        w.w_head, w.b_head = self.w_head.data_ptr(), self.b_head.data_ptr()
        self.w_we = f32(m.word_embed.weight); self.b_we = f32(m.word_embed.bias)
        self.lvl_pos = f32(m.lvl_embed.weight[m.lvl_1L[0]] + m.pos_1LC[0])              # (L, C)  var.py:164
        self.pos_start = f32(m.pos_start[0])                                              # (first_l, C)
        self.class_emb = f32(m.class_emb.weight)
        # one-pass attention needs the logit bound exp(min(scale_mul, ln 100)) <= 40 on every head (attention2_tcgen05.cu)
        w.attn_fixed_max = 0
        if m.attn_l2_norm:
            smax = max(float(blk.attn.scale_mul_1H11.detach().max()) for blk in m.blocks)
            w.attn_fixed_max = int(math.exp(min(smax, math.log(100.0))) <= 40.0)
        self.weights = w
        self._packed_key = key

    # ------------------------------------------------------------------ workspaces
    def begin(self, B: int, label_B: torch.Tensor, max_window_tokens: Optional[int] = None):
        """Allocate (or reuse) workspaces for batch B, reset the KV ring, compute the adaLN tables."""
        self.pack()
        m, dev = self.m, self.dev
        C, H, D, V = m.C, m.num_heads, m.depth, m.V
        imgs = 2 * B
        Lmax = m.L
        Lq_max = max_window_tokens or max(m.ls)
        key = (B, Lq_max)
        if key != self._ws_key:
            self.Lmax, self.Lmax_pad = Lmax, (Lmax + 7) // 8 * 8
            z = lambda *s, dt=torch.bfloat16: torch.zeros(*s, dtype=dt, device=dev)
            self.k_cache = [z(imgs, H, self.Lmax, 64) for _ in range(D)]
            self.v_cache = [z(imgs, H, 64, self.Lmax_pad) for _ in range(D)]
            Mx = imgs * Lq_max
            self.x = z(Mx, C, dt=torch.float32)
            self.xm, self.q, self.attn = z(Mx, C), z(Mx, C), z(Mx, C)
            self.hidden = z(Mx, 4 * C)
            self.logits = z(Mx, V, dt=torch.float32)
            self.cond_silu = z(imgs, C)
            self.ada = z(imgs, D * 6 * C, dt=torch.float32) if not m.shared_aln else z(D, imgs, 6 * C, dt=torch.float32)
            self.ada_shared = z(imgs, 6 * C, dt=torch.float32) if m.shared_aln else None
            self.head_mod = z(imgs, 2 * C, dt=torch.float32)
            self._ws_key = key
            self._graphs, self._graph_seen = {}, {}      # the captured passes point at the old buffers
        self.B, self.imgs, self.Lq_max = B, imgs, Lq_max
        self.kv_len = 0
        self._slot_keep = []      # slot maps of the sub-batch passes in flight (kept alive until the next begin())
        # cond = class_emb(cat(label, num_classes))  (models/var.py:162) -- an index gather, plumbing
        lab = label_B.to(dev)
        self.cond = self.class_emb[torch.cat((lab, torch.full_like(lab, m.num_classes)))].contiguous()
        _cabi.silu_bf16(self.cond, self.cond_silu)
        E = _cabi.GemmEpilogue
        if m.shared_aln:
            _cabi.gemm_bf16(self.cond_silu, C, self.w_ada, C, imgs, 6 * C, C,
                            E(epilogue=_cabi.EPI_F32, bias=self.b_ada.data_ptr(), out_f32=self.ada_shared.data_ptr(), ldo=6 * C))
            torch.add(self.ada_gss.view(D, 1, 6 * C), self.ada_shared.view(1, imgs, 6 * C), out=self.ada)   # basic_var.py:154
            self.ada_block_stride, self.ada_img_stride = imgs * 6 * C, 6 * C
        else:
            _cabi.gemm_bf16(self.cond_silu, C, self.w_ada, C, imgs, D * 6 * C, C,
                            E(epilogue=_cabi.EPI_F32, bias=self.b_ada.data_ptr(), out_f32=self.ada.data_ptr(), ldo=D * 6 * C))
            self.ada_block_stride, self.ada_img_stride = 6 * C, D * 6 * C
        _cabi.gemm_bf16(self.cond_silu, C, self.w_headnm, C, imgs, 2 * C, C,
                        E(epilogue=_cabi.EPI_F32, bias=self.b_headnm.data_ptr(), out_f32=self.head_mod.data_ptr(), ldo=2 * C))

    # ------------------------------------------------------------------ stage inputs
    def slot_map(self, slots: torch.Tensor) -> torch.Tensor:
        """int32 device map of a sub-batch pass: pass-image i < n is the conditional row of image slots[i], pass-image n+i its
        unconditional row (cache / adaLN slots slots[i] and B + slots[i])."""
        sl = slots.to(device=self.dev, dtype=torch.int32)
        mp = torch.cat((sl, sl + self.B)).contiguous()
        self._slot_keep.append(mp)
        return mp

    def put_first_map(self, Lq: int, tok_off: int = 0, slot_map: Optional[torch.Tensor] = None):
        m = self.m
        cond = self.cond if slot_map is None else self.cond[slot_map.long()].contiguous()
        _cabi.first_map(cond, cond.shape[0], m.first_l, m.C, self.pos_start, self.lvl_pos[:m.first_l], self.x, Lq, tok_off)

    def put_embed_map(self, si: int, next_map: torch.Tensor, Lq: int, tok_off: int = 0):
        m = self.m
        _cabi.embed_next_map(next_map, next_map.shape[0], m.ls[si], m.Cvae, m.C, self.w_we, self.b_we,
                             self.lvl_pos[m.begins[si]:m.ends[si]], self.x, Lq, tok_off)

    # ------------------------------------------------------------------ one pass
    def forward(self, stages: Sequence[int], want_logits: bool = True, check_position: bool = True,
                slot_map: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """Run the blocks (+head) over the inputs already placed in ``self.x`` for the consecutive ``stages``;
        appends their K/V at the cache position.  Returns logits viewed as (imgs, Lq, V).

        ``slot_map`` (from ``slot_map()``) runs the pass on a SUB-BATCH: its images keep their own cache / adaLN slots, the
        pass itself is dense (per-image ragged acceptance).  A sub-batch pass writes at the stage's own
        position ``begins[stages[0]]`` -- every image of the group has exactly its accepted prefix in the cache -- and does
        not move the engine-wide ``kv_len``."""
        m = self.m
        assert list(stages) == list(range(stages[0], stages[0] + len(stages)))
        imgs = self.imgs if slot_map is None else int(slot_map.shape[0])
        if slot_map is None:
            # token positions live in lvl_pos, not in the cache: sd_test3's target starts mid-pyramid with an empty cache
            assert not check_position or self.kv_len == m.begins[stages[0]], \
                f"KV cache holds {self.kv_len} tokens, stage {stages[0]} starts at {m.begins[stages[0]]}"
            kv_off = self.kv_len
        else:
            kv_off = m.begins[stages[0]]
        Lq = sum(m.ls[s] for s in stages)
        assert Lq <= self.Lq_max and imgs <= self.imgs
        p = _cabi.Pass()
        p.imgs, p.Lq, p.Lmax, p.Lmax_pad, p.kv_off, p.S = imgs, Lq, self.Lmax, self.Lmax_pad, kv_off, len(stages)
        p.slot_map = 0 if slot_map is None else slot_map.data_ptr()
        p.cache_slots = 0 if slot_map is None else self.imgs
        off = 0
        for j, s in enumerate(stages):
            p.seg_begin[j] = off
            off += m.ls[s]
        p.seg_begin[len(stages)] = off
        p.x, p.ada, p.head_mod = self.x.data_ptr(), self.ada.data_ptr(), self.head_mod.data_ptr()
        p.ada_block_stride, p.ada_img_stride = self.ada_block_stride, self.ada_img_stride
        for i in range(m.depth):
            p.k_cache[i] = self.k_cache[i].data_ptr()
            p.vT_cache[i] = self.v_cache[i].data_ptr()
        p.xm, p.q, p.attn, p.hidden = self.xm.data_ptr(), self.q.data_ptr(), self.attn.data_ptr(), self.hidden.data_ptr()
        p.logits = self.logits.data_ptr() if want_logits else 0
        gkey = (tuple(stages), kv_off, bool(want_logits), self._packed_key)
        if slot_map is not None or not self.use_graphs or _cabi.PROFILING or torch.cuda.is_current_stream_capturing():
            _cabi.var_forward(self.weights, p)
        elif gkey in self._graphs:
            g, n_launch = self._graphs[gkey]
            g.replay()
            _cabi.count_launches(n_launch)
        elif self._graph_seen.get(gkey, 0) == 0:
            self._graph_seen[gkey] = 1
            _cabi.var_forward(self.weights, p)          # first use: eager (also the warm-up every capture needs)
        else:
            g = torch.cuda.CUDAGraph()
            l0 = _cabi.launch_count()
            with torch.cuda.graph(g):
                _cabi.var_forward(self.weights, p)
            self._graphs[gkey] = (g, _cabi.launch_count() - l0)
            g.replay()                                   # capture records, it does not execute
        if slot_map is None:
            self.kv_len += Lq
        return self.logits[:imgs * Lq].view(imgs, Lq, m.V) if want_logits else None

    def kv_truncate(self, n: int):
        assert 0 <= n <= self.kv_len
        self.kv_len = n


class DeviceNoise:
    """Default noise provider: four device generators (draft / target / verify-u / resample streams), drawn in the
    order the loop spec fixes.  ``exponential`` is exactly the tensor torch.multinomial(n=1) draws."""

    STREAMS = ("draft", "target", "u", "resample")

    def __init__(self, seed: Optional[int], device):
        self.device = device
        self.g = {}
        for i, k in enumerate(self.STREAMS):
            if seed is None:
                self.g[k] = None
            else:
                self.g[k] = torch.Generator(device=device).manual_seed(seed * 4 + i)

    def exponential(self, stream: str, rows: int, V: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``out`` (a contiguous (rows, V) slice of a larger buffer) receives exactly the values a fresh tensor would."""
        t = torch.empty(rows, V, device=self.device, dtype=torch.float32) if out is None else out
        return t.exponential_(generator=self.g[stream])

    def uniform(self, stream: str, rows: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        t = torch.empty(rows, device=self.device, dtype=torch.float32) if out is None else out
        return t.uniform_(generator=self.g[stream])


class SingleGeneratorNoise:
    """One generator for everything -- the reference's convention in autoregressive_infer_cfg / sd_test3
    (models/var.py:144-145, 641-644): one Exp(1) draw of shape (B*l, V) per sampled stage."""

    def __init__(self, rng: Optional[torch.Generator], device):
        self.rng, self.device = rng, device

    def exponential(self, stream: str, rows: int, V: int) -> torch.Tensor:
        return torch.empty(rows, V, device=self.device, dtype=torch.float32).exponential_(generator=self.rng)

    def uniform(self, stream: str, rows: int) -> torch.Tensor:
        return torch.rand(rows, device=self.device, dtype=torch.float32, generator=self.rng)
