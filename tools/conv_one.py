#!/usr/bin/env python
"""one decoder-convolution launch shape (profiling target): python tools/conv_one.py [Cin] [Cout] [HW] [taps] [B]"""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdvar_b200 import _cabi
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 160
cout = int(sys.argv[2]) if len(sys.argv) > 2 else 160
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 256
taps = int(sys.argv[4]) if len(sys.argv) > 4 else 9
B = int(sys.argv[5]) if len(sys.argv) > 5 else 64
k = 3 if taps == 9 else 1
x = torch.randn(B, cin, hw, hw, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
wp = (torch.randn(taps, cout, cin, device="cuda") / math.sqrt(cin * taps)).bfloat16()
bias = torch.zeros(cout, device="cuda")
y = torch.empty(B, cout, hw, hw, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
for _ in range(4):
    _cabi.conv_nhwc(x, B, hw, hw, cin, wp, taps, cout, bias, None, y=y)
torch.cuda.synchronize()
print("ok")
