#!/usr/bin/env python
"""FID sampling CLI: python tools/sample_fid.py --out samples.npz [--vae vae.pth --draft var_d16.pth
--target var_d30.pth] [--mode sd|target] [--per-class 50] [--classes 1000] [--batch 100] [--png-dir DIR]

Without checkpoints (no network in this build) the models keep the deterministic hashed init of sdvar_b200.weights."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdvar_b200 import fid  # noqa: E402
from sdvar_b200.models import build_vae_var_speculative_decoding  # noqa: E402
from sdvar_b200.weights import var_state_dict, vqvae_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--vae"); ap.add_argument("--draft"); ap.add_argument("--target")
    ap.add_argument("--depth-draft", type=int, default=16); ap.add_argument("--depth-target", type=int, default=30)
    ap.add_argument("--mode", default="sd", choices=["sd", "target"])
    ap.add_argument("--per-class", type=int, default=50); ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=100); ap.add_argument("--gamma", type=int, default=2)
    ap.add_argument("--schedule", default="lockstep"); ap.add_argument("--png-dir")
    a = ap.parse_args()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    vae, d, t, sd = build_vae_var_speculative_decoding(dev, depth_draft=a.depth_draft, depth_target=a.depth_target)
    if a.vae and a.draft and a.target:
        fid.load_checkpoints(vae, a.vae, draft=(d, a.draft), target=(t, a.target))
    else:
        vae.load_state_dict(vqvae_state_dict(ch=160, device=dev))
        d.load_state_dict(var_state_dict(a.depth_draft, seed=1, tag="draft", device=dev))
        t.load_state_dict(var_state_dict(a.depth_target, seed=2, tag="target", device=dev))
    kw = dict(cfg=1.5, top_k=900, top_p=0.96)       # README.md:153
    if a.mode == "sd":
        gen = lambda B, lab, s: sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, g_seed=s, gamma=a.gamma, schedule=a.schedule, **kw)
    else:
        gen = lambda B, lab, s: t.autoregressive_infer_cfg(B, lab, g_seed=s, **kw)
    path = fid.sample_fid_set(gen, a.out, classes=range(a.classes), per_class=a.per_class, batch=a.batch, device=dev, png_dir=a.png_dir,
                              progress=lambda k, n: print(f"\r{k}/{n}", end="", file=sys.stderr))
    print("\nwrote", path)


if __name__ == "__main__":
    main()
