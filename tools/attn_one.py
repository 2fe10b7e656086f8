#!/usr/bin/env python
"""one attention launch shape (profiling target): python tools/attn_one.py [H] [stage]"""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdvar_b200 import _cabi
H = int(sys.argv[1]) if len(sys.argv) > 1 else 30
st = int(sys.argv[2]) if len(sys.argv) > 2 else 9
LS = [p * p for p in (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)]
SEG = [0] + [int(v) for v in np.cumsum(LS)]
imgs, Lq, kv_off = 128, LS[st], SEG[st]
q = torch.randn(imgs, H, Lq, 64, device="cuda").bfloat16()
kc = torch.nn.functional.normalize(torch.randn(imgs, H, 680, 64, device="cuda"), dim=-1).bfloat16()
vc = torch.randn(imgs, H, 64, 680, device="cuda").bfloat16()
o = torch.empty(imgs * Lq, H * 64, device="cuda", dtype=torch.bfloat16)
sm = torch.full((H,), math.log(4.0), device="cuda")
for _ in range(4):
    _cabi.attention(q, kc, vc, imgs, H, Lq, 680, 680, kv_off, [0, Lq], 1.0, o, logit_bound_log=sm)
torch.cuda.synchronize()
print("ok")
