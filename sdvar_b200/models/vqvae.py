"""VQVAE, inference side: quantizer (K5 kernels) + decoder.

The decoder (``fhat_to_img``, reference models/vqvae.py:62-63 + models/basic_vae.py:163-226) is NOT one of the
four north-star kernel families; it is the boundary right after the path.  It runs in bf16 channels-last:
every 3x3 / 1x1 convolution is the libsdvar tcgen05 implicit-GEMM kernel (``sdvar_conv_nhwc``: bias, skip connection and, for
``conv_out``, the clamp + fp32 NCHW image fused into its epilogue); GroupNorm+SiLU and the 2x upsampling are libsdvar kernels
too.  Shapes the kernel does not tile (input channels not a multiple of 32, widths that do not divide 128) and the encoder
fall back to cuDNN (library code).  Parameter names follow the reference checkpoint
(``decoder.*``, ``post_quant_conv.*``, ``quantize.*``); the encode side (``encoder.*``, ``quant_conv.*``) is
created when asked for or when a state dict that carries it (a real ``vae_ch160v4096z32.pth``) is loaded.
"""
from __future__ import annotations

from typing import Any, Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .quant import VectorQuantizer2


def _gn(c: int) -> nn.GroupNorm:
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)


_GN_SCRATCH = {}


def _fast(x: torch.Tensor) -> bool:
    """bf16 channels-last CUDA activations: the layout the libsdvar decoder kernels take."""
    return x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) \
        and x.shape[1] % 8 == 0


def _bias32(conv: nn.Conv2d) -> torch.Tensor:
    if getattr(conv, "_b32", None) is None or conv._b32.device != conv.bias.device:
        conv._b32 = conv.bias.detach().float().contiguous()
    return conv._b32


def _packed_w(conv: nn.Conv2d) -> torch.Tensor:
    """(taps, Cout, Cin) bf16: the K-major B operand of the implicit GEMM, one (Cout, Cin) matrix per filter tap"""
    w = conv.weight
    if getattr(conv, "_wp", None) is None or conv._wp.device != w.device or conv._wp_ver != w._version:
        co, ci, kh, kw = w.shape
        conv._wp = w.detach().permute(2, 3, 0, 1).reshape(kh * kw, co, ci).to(torch.bfloat16).contiguous()
        conv._wp_ver = w._version
    return conv._wp


def _packed_w_up(conv: nn.Conv2d) -> torch.Tensor:
    """(16, Cout, Cin) bf16 for conv3x3(nearest2x(x)): output pixel (2Y+a, 2X+b) reads the 2x2 input pixels (Y+a-1+u, X+b-1+v);
    matrix (a*2+b)*4 + u*2+v is the sum (in fp32, rounded once) of the 3x3 taps that land there: rows ky in R[a][u], columns kx
    in R[b][v] with R[0] = ({0}, {1,2}) and R[1] = ({0,1}, {2})"""
    w = conv.weight
    if getattr(conv, "_wpu", None) is None or conv._wpu.device != w.device or conv._wpu_ver != w._version:
        R = (((0,), (1, 2)), ((0, 1), (2,)))
        wf = w.detach().float()
        mats = []
        for a in range(2):
            for b in range(2):
                for u in range(2):
                    for v in range(2):
                        mats.append(sum(wf[:, :, ky, kx] for ky in R[a][u] for kx in R[b][v]))
        conv._wpu = torch.stack(mats).to(torch.bfloat16).contiguous()
        conv._wpu_ver = w._version
    return conv._wpu


def _tc_ok(conv: nn.Conv2d, x: torch.Tensor, nchw_f32: bool = False) -> bool:
    """shapes sdvar_conv_nhwc tiles (include/sdvar_b200.h): 3x3 padding 1 or 1x1, stride 1, Cin % 32 == 0, and 128 consecutive
    pixels form a box of the image"""
    if not _fast(x) or conv.stride != (1, 1) or conv.groups != 1 or conv.dilation != (1, 1):
        return False
    if not ((conv.kernel_size == (3, 3) and conv.padding == (1, 1)) or (conv.kernel_size == (1, 1) and conv.padding == (0, 0))):
        return False
    _, ci, H, W = x.shape
    if ci % 32 or (conv.out_channels % 8 and not nchw_f32):
        return False
    bw = min(W, 128)
    if W % bw or 128 % bw:
        return False
    bh = min(H, 128 // bw)
    return H % bh == 0 and (128 // bw) % bh == 0


def _conv_tc(conv: nn.Conv2d, x: torch.Tensor, bias: bool, res=None, image_out: bool = False) -> torch.Tensor:
    from .. import _cabi
    N, ci, H, W = x.shape
    co = conv.out_channels
    taps = conv.kernel_size[0] * conv.kernel_size[1]
    b = _bias32(conv) if bias and conv.bias is not None else None
    if image_out:
        y = torch.empty((N, co, H, W), device=x.device, dtype=torch.float32)
        _cabi.conv_nhwc(x, N, H, W, ci, _packed_w(conv), taps, co, b, None, y_f32_nchw=y, lo=-1.0, hi=1.0)
    else:
        y = torch.empty((N, co, H, W), device=x.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
        _cabi.conv_nhwc(x, N, H, W, ci, _packed_w(conv), taps, co, b, res, y=y)
    return y


def _conv_nobias(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    if _tc_ok(conv, x):
        return _conv_tc(conv, x, bias=False)
    return F.conv2d(x, conv.weight, None, conv.stride, conv.padding)


def _conv_bias_res(conv: nn.Conv2d, x: torch.Tensor, res=None) -> torch.Tensor:
    """conv(x) + bias (+ res).  Device path: the tcgen05 convolution with bias and skip connection in its epilogue.  Shapes it does
    not tile: bias-free cuDNN convolution, then ONE libsdvar pass adds bias and skip connection in place."""
    if _tc_ok(conv, x) and (res is None or (_fast(res) and res.shape[1] == conv.out_channels and res.shape[2:] == x.shape[2:])):
        return _conv_tc(conv, x, bias=True, res=res)
    if _fast(x) and conv.out_channels % 8 == 0:
        h = _conv_nobias(conv, x)
        if _fast(h) and (res is None or (_fast(res) and res.shape == h.shape)):
            from .. import _cabi
            N, C, H, W = h.shape
            _cabi.bias_residual_nhwc(h, _bias32(conv), res, N * H * W, C, h)
            return h
        h = h + conv.bias.view(1, -1, 1, 1)
        return h if res is None else res + h
    h = conv(x)
    return h if res is None else res + h


def _gn_act(norm: nn.GroupNorm, x: torch.Tensor, silu: bool, pre_bias=None) -> torch.Tensor:
    """GroupNorm (+SiLU) of x (+ pre_bias per channel).  On the device path (bf16, channels-last) this is the fused libsdvar kernel:
    PyTorch's GroupNorm round-trips channels-last bf16 through NCHW copies (measured 120 ms of a 140 ms decode of 64 images)."""
    if _fast(x) and x.shape[1] % 32 == 0 and norm.num_groups == 32:
        from .. import _cabi
        N, C, H, W = x.shape
        if getattr(norm, "_w32", None) is None or norm._w32.device != x.device:
            norm._w32, norm._b32 = norm.weight.detach().float().contiguous(), norm.bias.detach().float().contiguous()
        key = (x.device, N)
        if key not in _GN_SCRATCH:
            _GN_SCRATCH[key] = torch.empty(N * 128 * 64, device=x.device, dtype=torch.float32)
        y = torch.empty_like(x)
        _cabi.groupnorm_silu_nhwc(x, N, H * W, C, norm._w32, norm._b32, norm.eps, silu, y, _GN_SCRATCH[key], pre_bias=pre_bias)
        return y
    if pre_bias is not None:
        x = x + pre_bias.to(x.dtype).view(1, -1, 1, 1)
    y = norm(x)
    return F.silu(y) if silu else y


class _Res(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1, self.conv1 = _gn(cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2, self.conv2 = _gn(cout), nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        a = _gn_act(self.norm1, x, True)
        if _fast(a) and self.conv1.out_channels % 32 == 0:
            # conv1 runs bias-free; its bias is folded into norm2's statistics and affine
            b = _gn_act(self.norm2, _conv_nobias(self.conv1, a), True, pre_bias=_bias32(self.conv1))
        else:
            b = _gn_act(self.norm2, self.conv1(a), True)
        return _conv_bias_res(self.conv2, b, _conv_bias_res(self.nin_shortcut, x) if hasattr(self, "nin_shortcut") else x)


class _SpatialAttn(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.norm, self.qkv, self.proj_out = _gn(c), nn.Conv2d(c, 3 * c, 1), nn.Conv2d(c, c, 1)

    def forward(self, x):
        B, C, H, W = x.shape
        a = _gn_act(self.norm, x, False)
        if _tc_ok(self.qkv, a) and _tc_ok(self.proj_out, a):
            # channels-last all the way: qkv (B, HW, 3C) is read in place as three strided (B, HW, C) views
            t = _conv_bias_res(self.qkv, a).permute(0, 2, 3, 1).reshape(B, H * W, 3, C)
            o = F.scaled_dot_product_attention(t[:, None, :, 0], t[:, None, :, 1], t[:, None, :, 2], scale=C ** -0.5)
            o = o.reshape(B, H, W, C).permute(0, 3, 1, 2)          # a channels-last (B, C, H, W) view of the contiguous result
            return _conv_bias_res(self.proj_out, o, res=x)
        q, k, v = self.qkv(a).reshape(B, 3, C, H * W).unbind(1)
        o = F.scaled_dot_product_attention(q.transpose(1, 2).unsqueeze(1), k.transpose(1, 2).unsqueeze(1),
                                           v.transpose(1, 2).unsqueeze(1), scale=C ** -0.5)   # softmax(q k^T / sqrt(C)) v
        return x + self.proj_out(o.squeeze(1).transpose(1, 2).reshape(B, C, H, W))


class _Up(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        if _tc_ok(self.conv, x) and self.conv.kernel_size == (3, 3):
            # upsample + convolution in one step: four 2x2 parity convolutions on the low-resolution input (sdvar_conv_up2x_nhwc)
            from .. import _cabi
            N, C, H, W = x.shape
            co = self.conv.out_channels
            y = torch.empty((N, co, 2 * H, 2 * W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
            _cabi.conv_up2x_nhwc(x, N, H, W, C, _packed_w_up(self.conv), co, _bias32(self.conv), y)
            return y
        if _fast(x):
            from .. import _cabi
            N, C, H, W = x.shape
            y = torch.empty((N, C, 2 * H, 2 * W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
            _cabi.upsample2x_nhwc(x, N, H, W, C, y)
        else:
            y = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return _conv_bias_res(self.conv, y)


class _Level(nn.Module):
    def __init__(self, cin, cout, n, with_attn, with_up):
        super().__init__()
        self.block = nn.ModuleList([_Res(cin if i == 0 else cout, cout) for i in range(n)])
        self.attn = nn.ModuleList([_SpatialAttn(cout) for _ in range(n)] if with_attn else [])
        if with_up:
            self.upsample = _Up(cout)

    def forward(self, h):
        for i, b in enumerate(self.block):
            h = b(h)
            if len(self.attn):
                h = self.attn[i](h)
        return self.upsample(h) if hasattr(self, "upsample") else h


class _Mid(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.block_1, self.attn_1, self.block_2 = _Res(c, c), _SpatialAttn(c), _Res(c, c)

    def forward(self, h):
        return self.block_2(self.attn_1(self.block_1(h)))


class Decoder(nn.Module):
    def __init__(self, ch=160, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, in_channels=3, z_channels=32):
        super().__init__()
        n = len(ch_mult)
        c = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, c, 3, padding=1)
        self.mid = _Mid(c)
        levels = []
        for lv in reversed(range(n)):
            cout = ch * ch_mult[lv]
            levels.insert(0, _Level(c, cout, num_res_blocks + 1, with_attn=(lv == n - 1), with_up=(lv != 0)))
            c = cout
        self.up = nn.ModuleList(levels)
        self.norm_out, self.conv_out = _gn(c), nn.Conv2d(c, in_channels, 3, padding=1)

    def forward(self, z):
        h = self.mid(_conv_bias_res(self.conv_in, z))
        for lv in reversed(range(len(self.up))):
            h = self.up[lv](h)
        a = _gn_act(self.norm_out, h, True)
        if _tc_ok(self.conv_out, a, nchw_f32=True):
            return _conv_tc(self.conv_out, a, bias=True, image_out=True)      # fp32 NCHW, already clamped to [-1, 1]
        return self.conv_out(a)


class _Down(nn.Module):
    """Downsample2x (reference models/basic_vae.py:31-37): zero-pad right/bottom by one, 3x3 conv with stride 2."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, pad=(0, 1, 0, 1), mode="constant", value=0))


class _EncLevel(nn.Module):
    def __init__(self, cin, cout, n, with_attn, with_down):
        super().__init__()
        self.block = nn.ModuleList([_Res(cin if i == 0 else cout, cout) for i in range(n)])
        self.attn = nn.ModuleList([_SpatialAttn(cout) for _ in range(n)] if with_attn else [])
        if with_down:
            self.downsample = _Down(cout)

    def forward(self, h):
        for i, b in enumerate(self.block):
            h = b(h)
            if len(self.attn):
                h = self.attn[i](h)
        return self.downsample(h) if hasattr(self, "downsample") else h


class Encoder(nn.Module):
    """Encode side, reference models/basic_vae.py:99-160; same parameter names (``encoder.*``).  Runs in
    fp32 through PyTorch/cuDNN (library code, off the hot path): the nearest-code decision downstream is precision sensitive."""

    def __init__(self, ch=160, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, in_channels=3, z_channels=32):
        super().__init__()
        n = len(ch_mult)
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        in_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList([_EncLevel(ch * in_mult[lv], ch * ch_mult[lv], num_res_blocks, with_attn=(lv == n - 1),
                                             with_down=(lv != n - 1)) for lv in range(n)])
        c = ch * ch_mult[-1]
        self.mid = _Mid(c)
        self.norm_out, self.conv_out = _gn(c), nn.Conv2d(c, z_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for lv in self.down:
            h = lv(h)
        return self.conv_out(_gn_act(self.norm_out, self.mid(h), True))


class VQVAE(nn.Module):
    def __init__(self, vocab_size=4096, z_channels=32, ch=128, dropout=0.0, beta=0.25, using_znorm=False, quant_conv_ks=3,
                 quant_resi=0.5, share_quant_resi=4, default_qresi_counts=0, v_patch_nums=(1, 2, 3, 4, 5, 6, 8, 10, 13, 16),
                 test_mode=True, decoder_dtype=torch.bfloat16, with_encoder=False):
        super().__init__()
        self.test_mode, self.V, self.Cvae, self.vocab_size = test_mode, vocab_size, z_channels, vocab_size
        self._ch, self._qks = ch, quant_conv_ks
        self.decoder = Decoder(ch=ch, z_channels=z_channels)
        if with_encoder:
            self._build_encoder()
        self.downsample = 16
        self.quantize = VectorQuantizer2(vocab_size=vocab_size, Cvae=z_channels, using_znorm=using_znorm, beta=beta,
                                         default_qresi_counts=default_qresi_counts, v_patch_nums=v_patch_nums,
                                         quant_resi=quant_resi, share_quant_resi=share_quant_resi)
        self.post_quant_conv = nn.Conv2d(z_channels, z_channels, quant_conv_ks, padding=quant_conv_ks // 2)
        self.decoder_dtype = decoder_dtype
        self._dec_cache = None
        if test_mode:
            self.eval()
            for p in self.parameters():
                p.requires_grad_(False)

    def _build_encoder(self):
        """The encode side is optional: it is created on request or when a state dict carrying ``encoder.*`` is loaded."""
        if not hasattr(self, "encoder"):
            dev = self.post_quant_conv.weight.device if hasattr(self, "post_quant_conv") else None
            self.encoder = Encoder(ch=self._ch, z_channels=self.Cvae).to(dev)
            self.quant_conv = nn.Conv2d(self.Cvae, self.Cvae, self._qks, padding=self._qks // 2).to(dev)
            if self.test_mode:
                self.encoder.eval(); self.quant_conv.eval()
                for p in list(self.encoder.parameters()) + list(self.quant_conv.parameters()):
                    p.requires_grad_(False)

    def _decoder_exec(self):
        """bf16 channels-last copy of (post_quant_conv, decoder), refreshed when the fp32 master weights change."""
        ps = list(self.decoder.parameters()) + list(self.post_quant_conv.parameters())
        key = (str(ps[0].device), self.decoder_dtype, sum(p._version for p in ps))
        if self._dec_cache is None or self._dec_cache[0] != key:
            import copy
            mods = nn.Sequential(copy.deepcopy(self.post_quant_conv), copy.deepcopy(self.decoder))
            mods = mods.to(dtype=self.decoder_dtype).to(memory_format=torch.channels_last).eval()
            self._dec_cache = (key, mods)
        return self._dec_cache[1]

    @torch.no_grad()
    def fhat_to_img(self, f_hat: torch.Tensor) -> torch.Tensor:
        """decoder(post_quant_conv(f_hat)).clamp(-1,1)  (models/vqvae.py:62-63), fp32 result."""
        if self.decoder_dtype == torch.float32 or not f_hat.is_cuda:
            return self.decoder(self.post_quant_conv(f_hat.float())).clamp_(-1, 1)
        mods = self._decoder_exec()
        x = f_hat.to(self.decoder_dtype).contiguous(memory_format=torch.channels_last)
        y = _conv_bias_res(mods[0], x) if _tc_ok(mods[0], x) else mods[0](x)
        img = mods[1](y)
        return img if img.dtype == torch.float32 else img.float().clamp_(-1, 1)   # conv_out's epilogue already clamped

    def idxBl_to_img(self, ms_idx_Bl: List[torch.Tensor], same_shape: bool = True, last_one: bool = False):
        """models/vqvae.py:69-76: decode token pyramids (same_shape=True only)."""
        assert same_shape, "same_shape=False is the reference's experimental branch and not supported"
        B, HW = ms_idx_Bl[0].shape[0], self.quantize.v_patch_nums[-1]
        f_hat = torch.zeros(B, self.Cvae, HW, HW, device=ms_idx_Bl[0].device, dtype=torch.float32)
        outs = []
        for si, idx in enumerate(ms_idx_Bl):
            self.quantize.next_input_from_idx(si, f_hat, idx.contiguous())
            if not last_one:
                outs.append(self.fhat_to_img(f_hat))
        return self.fhat_to_img(f_hat) if last_one else outs

    def embed_to_img(self, ms_h_BChw: List[torch.Tensor], all_to_max_scale: bool = True, last_one: bool = False):
        """models/vqvae.py:78-82"""
        fh = self.quantize.embed_to_fhat(ms_h_BChw, all_to_max_scale=all_to_max_scale, last_one=last_one)
        return self.fhat_to_img(fh) if last_one else [self.fhat_to_img(f) for f in fh]

    @torch.no_grad()
    def encode_features(self, inp_img_no_grad: torch.Tensor) -> torch.Tensor:
        """quant_conv(encoder(img)) (models/vqvae.py:66), fp32 without TF32 convolutions."""
        if not hasattr(self, "encoder"):
            raise RuntimeError("this VQVAE was built without its encode side: construct it with with_encoder=True or load a "
                               "state dict that carries encoder.* / quant_conv.*")
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return self.quant_conv(self.encoder(inp_img_no_grad.float()))

    @torch.no_grad()
    def img_to_idxBl(self, inp_img_no_grad: torch.Tensor, v_patch_nums=None) -> List[torch.Tensor]:
        """models/vqvae.py:65-67: image (B,3,H,W) in [-1,1] -> token lists [(B, pn*pn) int64]."""
        return self.quantize.f_to_idxBl_or_fhat(self.encode_features(inp_img_no_grad), to_fhat=False, v_patch_nums=v_patch_nums)

    @torch.no_grad()
    def img_to_reconstructed_img(self, x: torch.Tensor, v_patch_nums=None, last_one: bool = False):
        """models/vqvae.py:84-90"""
        ls = self.quantize.f_to_idxBl_or_fhat(self.encode_features(x), to_fhat=True, v_patch_nums=v_patch_nums)
        return self.fhat_to_img(ls[-1]) if last_one else [self.fhat_to_img(f) for f in ls]

    def forward(self, *a, **k):
        raise NotImplementedError("VQVAE.forward is VAE training: out of scope")

    def load_state_dict(self, state_dict: Dict[str, Any], strict=True, assign=False):
        if any(k.startswith("encoder.") for k in state_dict):
            self._build_encoder()
        sd = dict(state_dict)
        if "quantize.ema_vocab_hit_SV" in sd and sd["quantize.ema_vocab_hit_SV"].shape != self.quantize.ema_vocab_hit_SV.shape:
            sd["quantize.ema_vocab_hit_SV"] = self.quantize.ema_vocab_hit_SV
        self._dec_cache = None
        return super().load_state_dict(sd, strict=strict, assign=assign)
