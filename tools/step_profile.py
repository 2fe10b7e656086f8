#!/usr/bin/env python
"""Where does a generation step go?  torch.profiler over one SD step at the bench config: CUDA time by kernel name + CPU-side gaps."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdvar_b200.models import build_vae_var_speculative_decoding
from sdvar_b200.weights import var_state_dict, vqvae_state_dict
P256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vae, d, t, sd = build_vae_var_speculative_decoding(dev, patch_nums=P256, depth_draft=16, depth_target=30)
vae.load_state_dict(vqvae_state_dict(ch=160, patch_nums=P256, device=dev))
d.load_state_dict(var_state_dict(16, patch_nums=P256, seed=1, tag="draft", device=dev))
t.load_state_dict(var_state_dict(30, patch_nums=P256, seed=2, tag="target", device=dev))
lab = torch.randint(0, 1000, (B,), device=dev)
run = lambda i: sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, g_seed=i, cfg=1.5, gamma=2, top_k=900, top_p=0.96)
for i in range(3): run(i)
torch.cuda.synchronize()
t0 = time.perf_counter(); run(5); torch.cuda.synchronize(); print("step wall ms", (time.perf_counter() - t0) * 1e3)
# decoder alone
f = torch.randn(B, 32, 16, 16, device=dev)
for _ in range(2): vae.fhat_to_img(f)
torch.cuda.synchronize(); t0 = time.perf_counter(); vae.fhat_to_img(f); torch.cuda.synchronize(); print("decoder ms", (time.perf_counter() - t0) * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(7); torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count, e.key) for e in ev), reverse=True)
tot = sum(r[0] for r in rows)
print("total device us", tot)
for us, n, k in rows[:28]: print(f"{us/1e3:9.2f} ms  n={n:5d}  {k[:100]}")
