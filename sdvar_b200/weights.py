"""Deterministic "random-init" recipe for VAR / VQVAE state dicts.

The reference leaves random init ill-defined: ``build_vae_var*`` turns
``reset_parameters`` into a no-op process-wide (reference models/__init__.py:31-32,
64-65), ``VAR.init_weights`` (models/var.py:261-311) re-initialises the transformer
but never reaches the VQVAE (hidden behind a tuple proxy, models/var.py:54-55), whose
53 conv tensors stay uninitialised memory unless ``vae_ch160v4096z32.pth`` is loaded.

This module pins one recipe that is *bit-reproducible on any machine and device*:
every tensor is filled from a splitmix64 integer hash of (seed, parameter name,
element index), mapped to a centred uniform value with exactly-rounded IEEE ops only
(int64 arithmetic, float64 mul/sub, one cast to float32).  The per-tensor standard
deviations follow ``VAR.init_weights`` (std = sqrt(1/(3C)), head x0.02, adaLN x0.5,
gamma rows x1e-5, proj/fc2 divided by sqrt(2*depth)); the distribution is uniform
instead of truncated normal, which is irrelevant to parity because the same state dict
is loaded into the reference (``load_state_dict(strict=True)``), the oracle and the
CUDA engine.

The key names and shapes are the reference's checkpoint surface, so a
real ``var_d*.pth`` / ``vae_ch160v4096z32.pth`` can be loaded instead.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch

_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    """two's-complement view of an unsigned 64-bit python int"""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _fnv1a(name: str) -> int:
    h = 0xCBF29CE484222325
    for b in name.encode():
        h = ((h ^ b) * 0x100000001B3) & _M64
    return h


def _lsr(z: torch.Tensor, n: int) -> torch.Tensor:
    """logical shift right on int64 tensors (torch's >> is arithmetic)"""
    return (z >> n) & ((1 << (64 - n)) - 1)


def hashed_uniform(name: str, seed: int, shape: Sequence[int], device="cpu") -> torch.Tensor:
    """float64 tensor in [-1, 1), element i = f(splitmix64(seed, fnv(name), i)); exact on CPU and CUDA."""
    n = int(math.prod(shape)) if len(shape) else 1
    base = _s64(_fnv1a(name) ^ ((seed * 0x9E3779B97F4A7C15) & _M64))
    z = torch.arange(n, dtype=torch.int64, device=device) * _s64(0x9E3779B97F4A7C15) + base
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    u = _lsr(z, 11).to(torch.float64) * (2.0 ** -53)  # [0,1), exact
    return (u * 2.0 - 1.0).reshape(tuple(shape))


def hashed(name: str, seed: int, shape: Sequence[int], std: float, device="cpu", dtype=torch.float32) -> torch.Tensor:
    """centred uniform with the requested std (uniform on [-a,a) has std a/sqrt(3))."""
    return (hashed_uniform(name, seed, shape, device) * (std * math.sqrt(3.0))).to(dtype)


def stage_lengths(patch_nums: Sequence[int]):
    return [pn * pn for pn in patch_nums]


def var_state_dict(depth: int, patch_nums: Sequence[int] = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16), V: int = 4096,
                   Cvae: int = 32, num_classes: int = 1000, shared_aln: bool = False, seed: int = 0,
                   device="cpu", embed_dim: int | None = None, num_heads: int | None = None,
                   init_adaln: float = 0.5, init_adaln_gamma: float = 1e-5, init_head: float = 0.02,
                   gamma_bias: float = 0.0, tag: str = "var") -> Dict[str, torch.Tensor]:
    """State dict with the reference VAR's keys/shapes (models/var.py:56-117, basic_var.py:58-150).

    ``gamma_bias`` (added to the two residual-gate rows of every adaLN) and ``init_head`` can be
    raised by tests to get active residual branches and peaked logits; with the defaults
    (models/__init__.py:24) the gates are ~1e-5 and the logits nearly uniform.
    """
    C = embed_dim or 64 * depth
    H = num_heads or depth
    L = sum(stage_lengths(patch_nums))
    K = len(patch_nums)
    std = math.sqrt(1.0 / C / 3.0)
    sd: Dict[str, torch.Tensor] = {}

    def P(name, shape, s):
        sd[name] = hashed(f"{tag}.{name}", seed, shape, s, device)

    P("pos_start", (1, patch_nums[0] ** 2, C), std)
    P("pos_1LC", (1, L, C), std)
    P("word_embed.weight", (C, Cvae), std)
    sd["word_embed.bias"] = torch.zeros(C, device=device)
    P("class_emb.weight", (num_classes + 1, C), std)
    P("lvl_embed.weight", (K, C), std)
    lvl = torch.cat([torch.full((pn * pn,), i, dtype=torch.int64) for i, pn in enumerate(patch_nums)]).view(1, L)
    sd["lvl_1L"] = lvl.to(device)
    d = lvl.view(1, L, 1)
    sd["attn_bias_for_masking"] = torch.where(d >= d.transpose(1, 2), 0.0, -torch.inf).reshape(1, 1, L, L).to(device)
    if shared_aln:
        P("shared_ada_lin.1.weight", (6 * C, C), std)
        sd["shared_ada_lin.1.bias"] = torch.zeros(6 * C, device=device)
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "attn.scale_mul_1H11"] = torch.full((1, H, 1, 1), math.log(4.0), device=device) \
            + hashed(f"{tag}.{p}scale_mul", seed, (1, H, 1, 1), 0.2, device)
        P(p + "attn.q_bias", (C,), 0.02)
        P(p + "attn.v_bias", (C,), 0.02)
        sd[p + "attn.zero_k_bias"] = torch.zeros(C, device=device)
        P(p + "attn.mat_qkv.weight", (3 * C, C), std)
        P(p + "attn.proj.weight", (C, C), std / math.sqrt(2 * depth))
        P(p + "attn.proj.bias", (C,), 0.02)
        P(p + "ffn.fc1.weight", (4 * C, C), std)
        P(p + "ffn.fc1.bias", (4 * C,), 0.02)
        P(p + "ffn.fc2.weight", (C, 4 * C), std / math.sqrt(2 * depth))
        P(p + "ffn.fc2.bias", (C,), 0.02)
        if shared_aln:
            g = hashed(f"{tag}.{p}ada_gss", seed, (1, 1, 6, C), 1.0 / math.sqrt(C), device)
            g[:, :, 2:] *= init_adaln
            g[:, :, :2] *= init_adaln_gamma
            g[:, :, :2] += gamma_bias
            sd[p + "ada_gss"] = g
        else:
            w = hashed(f"{tag}.{p}ada_lin.1.weight", seed, (6 * C, C), std, device)
            w[2 * C:] *= init_adaln
            w[:2 * C] *= init_adaln_gamma
            sd[p + "ada_lin.1.weight"] = w
            b = hashed(f"{tag}.{p}ada_lin.1.bias", seed, (6 * C,), 0.02, device)
            b[:2 * C] = b[:2 * C] * init_adaln_gamma + gamma_bias
            sd[p + "ada_lin.1.bias"] = b
    sd["head_nm.ada_lin.1.weight"] = hashed(f"{tag}.head_nm.w", seed, (2 * C, C), std * init_adaln, device)
    sd["head_nm.ada_lin.1.bias"] = torch.zeros(2 * C, device=device)
    sd["head.weight"] = hashed(f"{tag}.head.w", seed, (V, C), std * init_head, device)
    sd["head.bias"] = torch.zeros(V, device=device)
    return sd


def _decoder_shapes(ch: int, z: int, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, in_channels=3, prefix="decoder."):
    """(name, shape) list mirroring reference models/basic_vae.py:163-208 (Decoder.__init__)."""
    out = []

    def conv(name, cin, cout, k):
        out.append((name + ".weight", (cout, cin, k, k)))
        out.append((name + ".bias", (cout,)))

    def norm(name, c):
        out.append((name + ".weight", (c,)))
        out.append((name + ".bias", (c,)))

    def res(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    def attn(name, c):
        norm(name + ".norm", c); conv(name + ".qkv", c, 3 * c, 1); conv(name + ".proj_out", c, c, 1)

    nres = len(ch_mult)
    block_in = ch * ch_mult[nres - 1]
    conv(prefix + "conv_in", z, block_in, 3)
    res(prefix + "mid.block_1", block_in, block_in)
    attn(prefix + "mid.attn_1", block_in)
    res(prefix + "mid.block_2", block_in, block_in)
    for i_level in reversed(range(nres)):
        block_out = ch * ch_mult[i_level]
        for i_block in range(num_res_blocks + 1):
            res(f"{prefix}up.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
            if i_level == nres - 1:
                attn(f"{prefix}up.{i_level}.attn.{i_block}", block_in)
        if i_level != 0:
            conv(f"{prefix}up.{i_level}.upsample.conv", block_in, block_in, 3)
    norm(prefix + "norm_out", block_in)
    conv(prefix + "conv_out", block_in, in_channels, 3)
    return out


def _encoder_shapes(ch: int, z: int, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, in_channels=3, prefix="encoder."):
    """(name, shape) list mirroring reference models/basic_vae.py:99-142 (Encoder.__init__)."""
    out = []

    def conv(name, cin, cout, k):
        out.append((name + ".weight", (cout, cin, k, k)))
        out.append((name + ".bias", (cout,)))

    def norm(name, c):
        out.append((name + ".weight", (c,)))
        out.append((name + ".bias", (c,)))

    def res(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout, 3)
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    def attn(name, c):
        norm(name + ".norm", c); conv(name + ".qkv", c, 3 * c, 1); conv(name + ".proj_out", c, c, 1)

    nres = len(ch_mult)
    in_mult = (1,) + tuple(ch_mult)
    conv(prefix + "conv_in", in_channels, ch, 3)
    block_in = ch
    for i_level in range(nres):
        block_in, block_out = ch * in_mult[i_level], ch * ch_mult[i_level]
        for i_block in range(num_res_blocks):
            res(f"{prefix}down.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
            if i_level == nres - 1:
                attn(f"{prefix}down.{i_level}.attn.{i_block}", block_in)
        if i_level != nres - 1:
            conv(f"{prefix}down.{i_level}.downsample.conv", block_in, block_in, 3)
    res(prefix + "mid.block_1", block_in, block_in)
    attn(prefix + "mid.attn_1", block_in)
    res(prefix + "mid.block_2", block_in, block_in)
    norm(prefix + "norm_out", block_in)
    conv(prefix + "conv_out", block_in, z, 3)
    return out


def vqvae_state_dict(V: int = 4096, Cvae: int = 32, ch: int = 160, patch_nums: Sequence[int] = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16),
                     share_quant_resi: int = 4, seed: int = 0, device="cpu", with_encoder: bool = False,
                     tag: str = "vae") -> Dict[str, torch.Tensor]:
    """VQVAE state dict: quantizer + post_quant_conv + decoder, plus encoder + quant_conv with ``with_encoder``.

    Keys follow reference models/vqvae.py:36-50, models/basic_vae.py:99-208 and models/quant.py:27-39.
    """
    sd: Dict[str, torch.Tensor] = {}
    sd["quantize.embedding.weight"] = hashed(f"{tag}.codebook", seed, (V, Cvae), 1.0, device)
    sd["quantize.ema_vocab_hit_SV"] = torch.zeros(len(patch_nums), V, device=device)
    for i in range(share_quant_resi):
        sd[f"quantize.quant_resi.qresi_ls.{i}.weight"] = hashed(f"{tag}.phi{i}.w", seed, (Cvae, Cvae, 3, 3), 0.05, device)
        sd[f"quantize.quant_resi.qresi_ls.{i}.bias"] = hashed(f"{tag}.phi{i}.b", seed, (Cvae,), 0.02, device)
    sd["post_quant_conv.weight"] = hashed(f"{tag}.pqc.w", seed, (Cvae, Cvae, 3, 3), 0.06, device)
    sd["post_quant_conv.bias"] = hashed(f"{tag}.pqc.b", seed, (Cvae,), 0.02, device)
    shapes = _decoder_shapes(ch, Cvae)
    if with_encoder:
        shapes = shapes + _encoder_shapes(ch, Cvae) + [("quant_conv.weight", (Cvae, Cvae, 3, 3)), ("quant_conv.bias", (Cvae,))]
    for name, shape in shapes:
        if ".norm" in name and name.endswith(".weight"):
            sd[name] = 1.0 + hashed(f"{tag}.{name}", seed, shape, 0.05, device)
        elif name.endswith(".bias"):
            sd[name] = hashed(f"{tag}.{name}", seed, shape, 0.02, device)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = hashed(f"{tag}.{name}", seed, shape, 1.0 / math.sqrt(fan_in), device)
    return sd
