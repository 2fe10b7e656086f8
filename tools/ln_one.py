#!/usr/bin/env python
"""one ln_modulate launch shape (profiling target): python tools/ln_one.py [M] [C]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdvar_b200 import _cabi
M = int(sys.argv[1]) if len(sys.argv) > 1 else 54400
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
x = torch.randn(M, C, device="cuda")
mod = torch.randn(128, 6 * C, device="cuda")
o = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    _cabi.ln_modulate(x, M, C, M // 128, mod.data_ptr() + 8 * C, mod.data_ptr() + 16 * C, 6 * C, 1e-6, o)
torch.cuda.synchronize()
print("ok")
