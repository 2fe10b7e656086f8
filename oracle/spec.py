"""ctypes front-end of the plain-C arithmetic spec (oracle/spec_c/sdvar_spec.c).

TEST INFRASTRUCTURE ONLY.  Same argument meaning as the CUDA C-ABI entry points
``sdvar_sample_cfg_topk_topp`` / ``sdvar_verify_accept_resample`` (include/sdvar_b200.h),
minus the stream, on CPU numpy/torch buffers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "spec_c")])
    return os.path.join(_HERE, "_build", "libsdvar_spec.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libsdvar_spec.so")
        src = os.path.join(_HERE, "spec_c", "sdvar_spec.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            path = build()
        _LIB = C.CDLL(path)
        _LIB.sdvar_spec_expf.restype = C.c_float
        _LIB.sdvar_spec_expf.argtypes = [C.c_float]
    return _LIB


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def cfg_scalars(cfg: float, stage_ids: Sequence[int], K: int):
    """t1=fl32(1+t), t2=fl32(t), t=cfg*si/(K-1) in python double (models/var.py:190,199-200)."""
    t = [cfg * (si / (K - 1)) if K > 1 else 0.0 for si in stage_ids]
    return np.asarray([1 + x for x in t], np.float32), np.asarray(t, np.float32)


def top_p_threshold(top_p: float) -> float:
    """fl32(1 - top_p) or -1 when disabled (helpers.py:11-13)."""
    return float(np.float32(1.0 - top_p)) if top_p > 0 else -1.0


def expf(x: torch.Tensor) -> torch.Tensor:
    l = lib()
    return torch.tensor([l.sdvar_spec_expf(float(v)) for v in x.flatten().tolist()], dtype=torch.float32).view(x.shape)


def nearest_code(z_NC: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """Fixed-order nearest codebook entry (the arithmetic of sdvar_vq_nearest_code); z (N,C), codebook (V,C) fp32 CPU."""
    z = z_NC.detach().float().contiguous().cpu()
    e = codebook.detach().float().contiguous().cpu()
    out = torch.empty(z.shape[0], dtype=torch.int64)
    rc = lib().sdvar_spec_nearest_code(_p(z), _p(e), C.c_longlong(z.shape[0]), C.c_int(z.shape[1]), C.c_int(e.shape[0]), _p(out))
    assert rc == 0
    return out


def sample(logits_2BLV: torch.Tensor, seg_begin: Sequence[int], t1, t2, top_k: int, top_p: float,
           noise: Optional[torch.Tensor], want_mixed=True):
    """returns (idx (B,L) int64 or None, mixed (B,L,V) or None, prob (B,L) or None)"""
    x = logits_2BLV.contiguous().float()
    B2, L, V = x.shape
    B = B2 // 2
    S = len(seg_begin) - 1
    seg = np.asarray(seg_begin, np.int32)
    t1 = np.ascontiguousarray(t1, np.float32); t2 = np.ascontiguousarray(t2, np.float32)
    idx = torch.empty(B, L, dtype=torch.int64) if noise is not None else None
    prob = torch.empty(B, L, dtype=torch.float32) if noise is not None else None
    mixed = torch.empty(B, L, V, dtype=torch.float32) if want_mixed else None
    nz = None if noise is None else noise.contiguous().float()
    rc = lib().sdvar_spec_sample(_p(x), B, L, V, seg.ctypes.data_as(C.c_void_p), S, t1.ctypes.data_as(C.c_void_p),
                                 t2.ctypes.data_as(C.c_void_p), int(top_k), C.c_float(top_p_threshold(top_p)),
                                 _p(nz), _p(idx), _p(mixed), _p(prob))
    assert rc == 0
    return idx, mixed, prob


def verify(xt: torch.Tensor, xd: torch.Tensor, draft_idx: torch.Tensor, u: torch.Tensor, noise: torch.Tensor,
           seg_begin: Sequence[int]):
    """xt/xd (B,L,V) mixed+masked logits; returns dict of outputs (see sdvar_spec_verify)."""
    xt = xt.contiguous().float(); xd = xd.contiguous().float()
    B, L, V = xt.shape
    S = len(seg_begin) - 1
    seg = np.asarray(seg_begin, np.int32)
    out = dict(out_idx=torch.empty(B, L, dtype=torch.int64), accept=torch.empty(B, L, dtype=torch.uint8),
               p_d=torch.empty(B, L), q_d=torch.empty(B, L), first_reject=torch.empty(B, S, dtype=torch.int32),
               n_accept=torch.empty(B, S, dtype=torch.int32), accepted_stages=torch.empty(B, dtype=torch.int32),
               summary=torch.empty(4, dtype=torch.int32))
    rc = lib().sdvar_spec_verify(_p(xt), _p(xd), _p(draft_idx.contiguous().to(torch.int64)), _p(u.contiguous().float()),
                                 _p(noise.contiguous().float()), B, L, V, seg.ctypes.data_as(C.c_void_p), S,
                                 _p(out["out_idx"]), _p(out["accept"]), _p(out["p_d"]), _p(out["q_d"]),
                                 _p(out["first_reject"]), _p(out["n_accept"]), _p(out["accepted_stages"]), _p(out["summary"]))
    assert rc == 0
    return out
