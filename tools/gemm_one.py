#!/usr/bin/env python
"""one GEMM launch shape (profiling target): python tools/gemm_one.py [proj|fc2|qkv|fc1|head] [M] [C]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdvar_b200 import _cabi
kind = sys.argv[1] if len(sys.argv) > 1 else "proj"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
C = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
N, K, epi = {"qkv": (3 * C, C, "bf16"), "proj": (C, C, "resid"), "fc1": (4 * C, C, "gelu"), "fc2": (C, 4 * C, "resid"), "head": (4096, C, "f32")}[kind]
A = torch.randn(M, K, device="cuda").bfloat16()
W = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.zeros(N, device="cuda")
E = _cabi.GemmEpilogue
if epi == "f32":
    o = torch.empty(M, N, device="cuda"); e = E(epilogue=_cabi.EPI_F32, bias=bias.data_ptr(), out_f32=o.data_ptr(), ldo=N)
elif epi == "resid":
    o = torch.zeros(M, N, device="cuda"); gate = torch.ones(1, N, device="cuda")
    e = E(epilogue=_cabi.EPI_RESID_F32, bias=bias.data_ptr(), out_f32=o.data_ptr(), ldo=N, gate=gate.data_ptr(), ld_gate=N, tokens_per_img=M)
else:
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    e = E(epilogue=_cabi.EPI_GELU_BF16 if epi == "gelu" else _cabi.EPI_BF16, bias=bias.data_ptr(), out_bf16=o.data_ptr(), ldo=N)
for _ in range(4):
    _cabi.gemm_bf16(A, K, W, K, M, N, K, e)
torch.cuda.synchronize()
print("ok")
