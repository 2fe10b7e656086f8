"""ctypes binding of libsdvar_b200.so (include/sdvar_b200.h).

This is the only way the package reaches device code: there is no PyTorch/CPU fallback.  ``lib()``
raises if the shared library is missing; every compute entry returns SDVAR_ERR_ARCH on a non-sm_100
device and the wrappers below turn any non-zero status into ``SdvarError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsdvar_b200.so")
MAX_SEG = 16
MAX_DEPTH = 64

EPI_F32, EPI_BF16, EPI_GELU_BF16, EPI_RESID_F32, EPI_QKV = range(5)

EXPORTS = [
    "sdvar_abi_version", "sdvar_last_error", "sdvar_arch_check", "sdvar_num_sms", "sdvar_launch_count", "sdvar_count_launches",
    "sdvar_sample_cfg_topk_topp", "sdvar_verify_accept_resample", "sdvar_verify_workspace_bytes", "sdvar_verify_top1", "sdvar_vq_next_input", "sdvar_vq_area_down",
    "sdvar_embed_next_map", "sdvar_first_map", "sdvar_ln_modulate", "sdvar_silu_bf16", "sdvar_f32_to_bf16", "sdvar_image_to_u8",
    "sdvar_gemm_bf16", "sdvar_attention", "sdvar_var_forward", "sdvar_profile_begin", "sdvar_profile_end",
    "sdvar_groupnorm_silu_nhwc", "sdvar_conv_nhwc", "sdvar_conv_up2x_nhwc", "sdvar_bias_residual_nhwc", "sdvar_upsample2x_nhwc", "sdvar_vq_nearest_code",
    "sdvar_debug_spec_expf",
]
PROFILE_FAMILIES = ("gemm", "attention", "ln_modulate", "sample", "verify", "vq", "embed", "misc", "conv")


class SdvarError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [("epilogue", C.c_int), ("bias", C.c_void_p), ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p),
                ("ldo", C.c_int), ("gate", C.c_void_p), ("ld_gate", C.c_int), ("tokens_per_img", C.c_int), ("slot_map", C.c_void_p),
                ("q_out", C.c_void_p), ("k_cache", C.c_void_p), ("vT_cache", C.c_void_p), ("scale_mul", C.c_void_p),
                ("H", C.c_int), ("Lq", C.c_int), ("Lmax", C.c_int), ("Lmax_pad", C.c_int), ("kv_off", C.c_int),
                ("l2norm", C.c_int)]


class VarWeights(C.Structure):
    _fields_ = [("depth", C.c_int), ("C", C.c_int), ("H", C.c_int), ("V", C.c_int), ("Cvae", C.c_int), ("l2norm", C.c_int),
                ("eps", C.c_float), ("attn_scale", C.c_float),
                ("w_qkv", C.c_void_p * MAX_DEPTH), ("b_qkv", C.c_void_p * MAX_DEPTH), ("scale_mul", C.c_void_p * MAX_DEPTH),
                ("w_proj", C.c_void_p * MAX_DEPTH), ("b_proj", C.c_void_p * MAX_DEPTH),
                ("w_fc1", C.c_void_p * MAX_DEPTH), ("b_fc1", C.c_void_p * MAX_DEPTH),
                ("w_fc2", C.c_void_p * MAX_DEPTH), ("b_fc2", C.c_void_p * MAX_DEPTH),
                ("w_head", C.c_void_p), ("b_head", C.c_void_p), ("attn_fixed_max", C.c_int)]


class Pass(C.Structure):
    _fields_ = [("imgs", C.c_int), ("Lq", C.c_int), ("Lmax", C.c_int), ("Lmax_pad", C.c_int), ("kv_off", C.c_int),
                ("S", C.c_int), ("seg_begin", C.c_int * (MAX_SEG + 1)), ("slot_map", C.c_void_p), ("cache_slots", C.c_int),
                ("x", C.c_void_p), ("ada", C.c_void_p), ("ada_block_stride", C.c_longlong), ("ada_img_stride", C.c_longlong),
                ("head_mod", C.c_void_p),
                ("k_cache", C.c_void_p * MAX_DEPTH), ("vT_cache", C.c_void_p * MAX_DEPTH),
                ("xm", C.c_void_p), ("q", C.c_void_p), ("attn", C.c_void_p), ("hidden", C.c_void_p), ("logits", C.c_void_p)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SdvarError(f"{LIB_PATH} is missing: build it with `python -m sdvar_b200.build` "
                             f"(there is no CPU / PyTorch fallback for the SDVAR hot path)")
        l = C.CDLL(LIB_PATH)
        l.sdvar_last_error.restype = C.c_char_p
        l.sdvar_launch_count.restype = C.c_longlong
        l.sdvar_verify_workspace_bytes.restype = C.c_longlong
        _LIB = l
    return _LIB


def _check(rc: int, what: str):
    if rc != 0:
        raise SdvarError(f"{what} failed ({rc}): {lib().sdvar_last_error().decode()}")


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise SdvarError("C-ABI arguments must be CUDA tensors: libsdvar_b200 has no CPU path")
    assert t.is_contiguous(), "C-ABI arguments must be contiguous"
    if t.device.index != torch.cuda.current_device():
        # the library launches on the CURRENT device and stream: a pointer of another GPU would be dereferenced there
        raise SdvarError(f"tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                         f"wrap the call in torch.cuda.device({t.device.index})")
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _iarr(v: Sequence[int]):
    return (C.c_int * len(v))(*[int(x) for x in v])


def _farr(v: Sequence[float]):
    return (C.c_float * len(v))(*[float(x) for x in v])


def launch_count() -> int:
    return int(lib().sdvar_launch_count())


# ---- thin typed wrappers (argument meaning: include/sdvar_b200.h) ---------------------------------
def sample_cfg_topk_topp(logits_2BLV, B, L, V, seg_begin, t1, t2, top_k, one_minus_top_p, noise, idx_out, mixed_out,
                         prob_out, in_ld=None, in_off=0, out_ld=None, out_off=0):
    _check(lib().sdvar_sample_cfg_topk_topp(ptr(logits_2BLV), B, L, L if in_ld is None else in_ld, in_off,
                                            L if out_ld is None else out_ld, out_off, V, _iarr(seg_begin), len(seg_begin) - 1, _farr(t1),
                                            _farr(t2), int(top_k), C.c_float(one_minus_top_p), ptr(noise), ptr(idx_out),
                                            ptr(mixed_out), ptr(prob_out), stream_ptr()), "sdvar_sample_cfg_topk_topp")


def verify_workspace_ints(B, S) -> int:
    """number of int32 elements the verify workspace needs (sdvar_verify_workspace_bytes / 4); zero it once"""
    n = int(lib().sdvar_verify_workspace_bytes(int(B), int(S)))
    if n < 0:
        raise SdvarError(f"sdvar_verify_workspace_bytes({B},{S}) failed ({n})")
    return n // 4


def verify_accept_resample(xt, xd, draft_idx, u, noise, B, L, V, seg_begin, out_idx, accept, p_d, q_d, first_reject,
                           n_accept, accepted_stages, summary, workspace, stage_major_aux=False):
    S = len(seg_begin) - 1
    if workspace.numel() < verify_workspace_ints(B, S):
        raise SdvarError(f"verify workspace too small: {workspace.numel()} int32 < {verify_workspace_ints(B, S)}")
    _check(lib().sdvar_verify_accept_resample(ptr(xt), ptr(xd), ptr(draft_idx), ptr(u), ptr(noise), int(bool(stage_major_aux)), B, L, V,
                                              _iarr(seg_begin), S, ptr(out_idx), ptr(accept), ptr(p_d),
                                              ptr(q_d), ptr(first_reject), ptr(n_accept), ptr(accepted_stages),
                                              ptr(summary), ptr(workspace), stream_ptr()), "sdvar_verify_accept_resample")


def verify_top1(xt, draft_idx, B, L, V, seg_begin, match, n_match):
    _check(lib().sdvar_verify_top1(ptr(xt), ptr(draft_idx), B, L, V, _iarr(seg_begin), len(seg_begin) - 1, ptr(match),
                                   ptr(n_match), stream_ptr()), "sdvar_verify_top1")


def vq_next_input(idx_Bl, B, pn, HW, pn_next, Cvae, codebook, phi_w, phi_b, f_hat, next_map, resi_ratio=0.5, f_rest=None):
    _check(lib().sdvar_vq_next_input(ptr(idx_Bl), B, pn, HW, pn_next, Cvae, ptr(codebook), ptr(phi_w), ptr(phi_b),
                                     C.c_float(resi_ratio), ptr(f_hat), ptr(next_map), ptr(f_rest), stream_ptr()), "sdvar_vq_next_input")


def vq_area_down(f_hat, B, HW, pn_next, Cvae, next_map):
    _check(lib().sdvar_vq_area_down(ptr(f_hat), B, HW, pn_next, Cvae, ptr(next_map), stream_ptr()), "sdvar_vq_area_down")


def vq_nearest_code(z_NC, codebook, N, Cvae, V, idx_out):
    _check(lib().sdvar_vq_nearest_code(ptr(z_NC), ptr(codebook), C.c_longlong(N), Cvae, V, ptr(idx_out), stream_ptr()),
           "sdvar_vq_nearest_code")


def debug_spec_expf(x, y_packed, y_scalar):
    _check(lib().sdvar_debug_spec_expf(ptr(x), C.c_longlong(x.numel()), ptr(y_packed), ptr(y_scalar), stream_ptr()),
           "sdvar_debug_spec_expf")


def embed_next_map(next_map, B, l, Cvae, Cm, W, b, lvl_pos, x, ldx_tokens, tok_off):
    _check(lib().sdvar_embed_next_map(ptr(next_map), B, l, Cvae, Cm, ptr(W), ptr(b), ptr(lvl_pos), ptr(x), ldx_tokens,
                                      tok_off, stream_ptr()), "sdvar_embed_next_map")


def first_map(cond, B2, first_l, Cm, pos_start, lvl_pos, x, ldx_tokens, tok_off):
    _check(lib().sdvar_first_map(ptr(cond), B2, first_l, Cm, ptr(pos_start), ptr(lvl_pos), ptr(x), ldx_tokens, tok_off,
                                 stream_ptr()), "sdvar_first_map")


def ln_modulate(x, M, Cm, tokens_per_img, scale, shift, ld_mod, eps, out, slot_map=None):
    _check(lib().sdvar_ln_modulate(ptr(x), M, Cm, tokens_per_img, C.c_void_p(scale), C.c_void_p(shift), ld_mod, ptr(slot_map),
                                   C.c_float(eps), ptr(out), stream_ptr()), "sdvar_ln_modulate")


def silu_bf16(x, out):
    _check(lib().sdvar_silu_bf16(ptr(x), C.c_longlong(x.numel()), ptr(out), stream_ptr()), "sdvar_silu_bf16")


def f32_to_bf16(x, out):
    _check(lib().sdvar_f32_to_bf16(ptr(x), C.c_longlong(x.numel()), ptr(out), stream_ptr()), "sdvar_f32_to_bf16")


def image_to_u8(img_B3HW, out, hwc=False):
    B, _, H, W = img_B3HW.shape
    _check(lib().sdvar_image_to_u8(ptr(img_B3HW), B, H, W, int(bool(hwc)), ptr(out), stream_ptr()), "sdvar_image_to_u8")


def gemm_bf16(A, lda, W, ldw, M, N, K, epi: GemmEpilogue):
    _check(lib().sdvar_gemm_bf16(ptr(A), lda, ptr(W), ldw, M, N, K, C.byref(epi), stream_ptr()), "sdvar_gemm_bf16")


def attention(q, k_cache, vT_cache, imgs, H, Lq, Lmax, Lmax_pad, kv_off, seg_begin, scale, out, logit_bound_log=None,
              slot_map=None, cache_slots=0):
    _check(lib().sdvar_attention(ptr(q), ptr(k_cache), ptr(vT_cache), imgs, H, Lq, Lmax, Lmax_pad, kv_off,
                                 _iarr(seg_begin), len(seg_begin) - 1, C.c_float(scale), ptr(logit_bound_log), ptr(slot_map),
                                 int(cache_slots), ptr(out), stream_ptr()),
           "sdvar_attention")


def var_forward(w: VarWeights, p: Pass):
    _check(lib().sdvar_var_forward(C.byref(w), C.byref(p), stream_ptr()), "sdvar_var_forward")


PROFILING = False     # while on, engines launch eagerly (the per-family event brackets cannot live inside a CUDA graph)


def count_launches(n: int):
    lib().sdvar_count_launches(C.c_longlong(n))


def profile_begin():
    global PROFILING
    _check(lib().sdvar_profile_begin(), "sdvar_profile_begin")
    PROFILING = True


def profile_end() -> dict:
    """{family: (ms, algorithmic work [FLOP for gemm/attention, bytes otherwise], launches)}"""
    global PROFILING
    PROFILING = False
    n = len(PROFILE_FAMILIES)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
    _check(lib().sdvar_profile_end(ms, work, cnt), "sdvar_profile_end")
    return {PROFILE_FAMILIES[i]: (ms[i], work[i], int(cnt[i])) for i in range(n)}


def groupnorm_silu_nhwc(x, N, HW, Cc, gamma, beta, eps, silu, y, scratch, pre_bias=None):
    _check(lib().sdvar_groupnorm_silu_nhwc(C.c_void_p(x.data_ptr()), ptr(pre_bias), N, HW, Cc, ptr(gamma), ptr(beta), C.c_float(eps),
                                           int(silu), C.c_void_p(y.data_ptr()), ptr(scratch), stream_ptr()), "sdvar_groupnorm_silu_nhwc")


def bias_residual_nhwc(h, bias, res, rows, Cc, out):
    _check(lib().sdvar_bias_residual_nhwc(C.c_void_p(h.data_ptr()), ptr(bias), C.c_void_p(res.data_ptr() if res is not None else 0), C.c_longlong(rows), Cc,
                                          C.c_void_p(out.data_ptr()), stream_ptr()), "sdvar_bias_residual_nhwc")


def conv_nhwc(x, N, H, W, Cin, w_packed, taps, Cout, bias, res, y=None, y_f32_nchw=None, lo=-1.0, hi=1.0):
    """tcgen05 implicit-GEMM convolution (3x3 padding 1, or 1x1) on channels-last bf16; see include/sdvar_b200.h"""
    vp = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
    _check(lib().sdvar_conv_nhwc(vp(x), N, H, W, Cin, vp(w_packed), taps, Cout, ptr(bias), vp(res), vp(y), ptr(y_f32_nchw),
                                 C.c_float(lo), C.c_float(hi), stream_ptr()), "sdvar_conv_nhwc")


def conv_up2x_nhwc(x, N, H, W, Cin, w_par, Cout, bias, y):
    """conv3x3(nearest2x(x)) + bias as four 2x2 parity convolutions on the low-resolution input; see include/sdvar_b200.h"""
    _check(lib().sdvar_conv_up2x_nhwc(C.c_void_p(x.data_ptr()), N, H, W, Cin, C.c_void_p(w_par.data_ptr()), Cout, ptr(bias),
                                      C.c_void_p(y.data_ptr()), stream_ptr()), "sdvar_conv_up2x_nhwc")


def upsample2x_nhwc(x, N, H, W, Cc, y):
    _check(lib().sdvar_upsample2x_nhwc(C.c_void_p(x.data_ptr()), N, H, W, Cc, C.c_void_p(y.data_ptr()), stream_ptr()),
           "sdvar_upsample2x_nhwc")
