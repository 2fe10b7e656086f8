#!/usr/bin/env python
"""one K3 launch shape (profiling target): python tools/sample_one.py [B] [l] [top_k] [top_p] [mixed]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdvar_b200 import _cabi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
l = int(sys.argv[2]) if len(sys.argv) > 2 else 256
tk = int(sys.argv[3]) if len(sys.argv) > 3 else 900
tp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.96
want_mixed = (sys.argv[5] != "0") if len(sys.argv) > 5 else True
V = 4096
lg = torch.randn(2 * B, l, V, device="cuda")
noise = torch.empty(B * l, V, device="cuda").exponential_()
idx = torch.empty(B, l, dtype=torch.int64, device="cuda")
mixed = torch.empty(B, l, V, device="cuda") if want_mixed else None
thr = float(np.float32(1 - tp)) if tp > 0 else -1.0
for _ in range(4):
    _cabi.sample_cfg_topk_topp(lg, B, l, V, [0, l], [1.5], [0.5], tk, thr, noise, idx, mixed, None)
torch.cuda.synchronize()
print("ok", int(idx.sum()))
