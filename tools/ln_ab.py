import os, sys, torch
sys.path.insert(0, os.getcwd())
from sdvar_b200 import _cabi
DEV="cuda"
def timeit(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
big = torch.empty(256*1024*1024//4, device=DEV)
for C in (1920,1024):
    for l in (36,64,100,169,256):
        M=128*l
        x=torch.randn(M,C,device=DEV); mod=torch.randn(128,6*C,device=DEV); o=torch.empty(M,C,device=DEV,dtype=torch.bfloat16)
        ms=timeit(lambda:_cabi.ln_modulate(x,M,C,l,mod.data_ptr()+8*C,mod.data_ptr()+16*C,6*C,1e-6,o))
        print(f"C={C} M={M} tpi={l} {ms*1000:7.1f} us {M*C*6/ms/1e6:6.0f} GB/s")
