#!/usr/bin/env python
"""Sweep of the draft-then-verify loop on one B200: batch sizes 1/3/17/100 x gamma 1/2/4 x three filter settings x both accept
rules on small models (d4 -> d6, 256 px pyramid); checks run-to-run determinism (bit-identical tokens and images) and the
acceptance bookkeeping.  `python tools/fuzz_sd.py` prints `bad 0` when everything holds."""
import sys, os, itertools
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from sdvar_b200.models import build_vae_var_speculative_decoding
from sdvar_b200.weights import var_state_dict, vqvae_state_dict
DEV="cuda"
P256=(1,2,3,4,5,6,8,10,13,16)
LS=[p*p for p in P256]
vae, draft, target, sd = build_vae_var_speculative_decoding(device=DEV, patch_nums=P256, depth_draft=4, depth_target=6, ch=32)
vae.load_state_dict(vqvae_state_dict(ch=32, patch_nums=P256, device=DEV))
draft.load_state_dict(var_state_dict(4, patch_nums=P256, seed=1, tag="draft", device=DEV))
target.load_state_dict(var_state_dict(6, patch_nums=P256, seed=2, tag="target", device=DEV))
bad=0
for B,gamma,(tk,tp),rule in itertools.product((1,3,17,100),(1,2,4),((0,0.0),(900,0.96),(50,0.5)),("speculative","reference")):
    labels=torch.randint(0,1000,(B,),generator=torch.Generator().manual_seed(B)).to(DEV)
    outs=[]
    for rep in range(2):
        img,toks,fh=sd.sdvar_autoregressive_infer_cfg_parallel_v1(B=B,label_B=labels,g_seed=7,cfg=1.5,gamma=gamma,top_k=tk,top_p=tp,accept_rule=rule,return_tokens=True)
        torch.cuda.synchronize(); outs.append((img.clone(),[t.clone() for t in toks],dict(sd.last_stats)))
    (i0,t0,s0),(i1,t1,s1)=outs
    ok = torch.equal(i0,i1) and all(torch.equal(a,b) for a,b in zip(t0,t1)) and s0==s1
    ok = ok and [tuple(t.shape) for t in t0]==[(B,l) for l in LS] and bool(torch.isfinite(i0).all()) and sum(s0["advance"])==10
    ok = ok and all(int(t.min())>=0 and int(t.max())<4096 for t in t0)
    if rule=="speculative": ok = ok and s0["accepted_tokens"]+s0["rejected_tokens"]==B*680
    if not ok: bad+=1; print("FAIL",B,gamma,tk,tp,rule,s0)
# baseline loop + sd_test3 identity for odd batch
img=target.autoregressive_infer_cfg(B=5,label_B=3,g_seed=1,cfg=1.5,top_k=900,top_p=0.96)
img2=sd.sdvar_autoregressive_infer_cfg_sd_test3(B=5,label_B=3,g_seed=1,cfg=1.5,top_k=900,top_p=0.96,entry_num=0)
print("sd_test3(entry 0)==target:", torch.equal(img,img2))
print("bad",bad)
