#!/usr/bin/env python
"""one K4 verify launch shape (profiling target): python tools/verify_one.py [B] [scale]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdvar_b200 import _cabi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
L, V, S = 680, 4096, 10
SEG = [0] + [int(v) for v in np.cumsum([p * p for p in (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)])]
g = torch.Generator(device="cuda").manual_seed(0)
xt = torch.randn(B, L, V, device="cuda", generator=g) * scale
xd = torch.randn(B, L, V, device="cuda", generator=g) * scale
d = torch.multinomial(xd.view(-1, V).softmax(-1), 1).view(B, L)
u = torch.rand(B, L, device="cuda")
noise = torch.empty(B * L, V, device="cuda").exponential_()
o = torch.empty(B, L, dtype=torch.int64, device="cuda"); a = torch.empty(B, L, dtype=torch.uint8, device="cuda")
fr = torch.empty(B, S, dtype=torch.int32, device="cuda"); na = torch.empty(B, S, dtype=torch.int32, device="cuda")
st = torch.empty(B, dtype=torch.int32, device="cuda"); sm = torch.empty(4, dtype=torch.int32, device="cuda")
ws = torch.zeros(_cabi.verify_workspace_ints(B, S), dtype=torch.int32, device="cuda")
for _ in range(4):
    _cabi.verify_accept_resample(xt, xd, d, u, noise, B, L, V, SEG, o, a, None, None, fr, na, st, sm, ws)
torch.cuda.synchronize()
print("ok", sm.tolist())
