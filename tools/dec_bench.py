#!/usr/bin/env python
"""Decoder (fhat_to_img, B=64, ch=160, 256 px) timing on one B200: whole call + per-family CUDA-event split.

    python tools/dec_bench.py [--batch 64] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdvar_b200 import _cabi  # noqa: E402
from sdvar_b200.models.vqvae import VQVAE  # noqa: E402
from sdvar_b200.weights import hashed, vqvae_state_dict  # noqa: E402

P256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda")
    vae = VQVAE(vocab_size=4096, z_channels=32, ch=160, v_patch_nums=P256).to(dev)
    vae.load_state_dict(vqvae_state_dict(ch=160, patch_nums=P256, device=dev))
    f_hat = hashed("decbench", 0, (a.batch, 32, 16, 16), 1.0).to(dev)
    for _ in range(3):
        img = vae.fhat_to_img(f_hat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        img = vae.fhat_to_img(f_hat)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    _cabi.profile_begin()
    vae.fhat_to_img(f_hat)
    prof = _cabi.profile_end()
    fl = 0.0
    out = {"batch": a.batch, "ms_per_decode": ms, "images_per_s": a.batch / ms * 1e3, "families": {}}
    for fam, (fms, work, cnt) in prof.items():
        if cnt:
            out["families"][fam] = {"ms": fms, "launches": cnt, "rate": work / (fms * 1e-3) / (1e12 if fam == "conv" else 1e9),
                                    "unit": "TFLOP/s" if fam == "conv" else "GB/s"}
    print(json.dumps(out, indent=1))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    assert bool(torch.isfinite(img).all())


if __name__ == "__main__":
    main()
