#!/usr/bin/env python
"""Flip rate of the C arithmetic spec of K3 (oracle/spec_c: sdvar_spec_sample) against the REFERENCE sampler
(/root/reference/models/helpers.py:6-19, sample_with_top_k_top_p_) on identical logits and identical noise.

The spec fixes an exp polynomial and (on the filtered path) fixed-point sums where the reference leaves the evaluation
order to ATen, so a token can differ on a near-tie of the exponential race or of the top-p cut.  This tool measures how
often: rows of N(0, scale^2) cond/uncond logits, CFG-mixed as models/var.py:199-200 does, sampled by both with the noise
torch.multinomial(n=1) draws (pin P4: multinomial == argmax(p / Exp(1)) with the same generator).

    python tools/flip_rate.py [--rows-per-config 262144] [--out profiles/flip_rate_r02.json]

Runs in the build container only (needs /root/reference); the committed JSON is what tests/test_flip_rate.py asserts on,
next to a small live sample when the reference is present.  TEST INFRASTRUCTURE (imports oracle/)."""
import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CONFIGS = [(0.05, 0, 0.0), (3.0, 0, 0.0), (0.05, 900, 0.96), (3.0, 900, 0.96)]
V, L = 4096, 64


def _ref_sampler():
    sys.path.insert(0, "/root/reference")
    with contextlib.redirect_stdout(io.StringIO()):
        from models.helpers import sample_with_top_k_top_p_
    return sample_with_top_k_top_p_


def chunk(args):
    ci, seed, B = args
    from oracle import spec
    torch.set_num_threads(1)
    sampler = _ref_sampler()
    scale, tk, tp = CONFIGS[ci]
    g = torch.Generator().manual_seed(1_000_003 * ci + seed)
    lg = torch.randn(2 * B, L, V, generator=g) * scale
    si, K, cfg = 6, 10, 1.5
    t = cfg * (si / (K - 1))
    mixed = (1 + t) * lg[:B] - t * lg[B:]
    state = g.get_state()
    ref = sampler(mixed.clone(), rng=g, top_k=tk, top_p=tp, num_samples=1)[:, :, 0]
    g.set_state(state)
    noise = torch.empty(B * L, V).exponential_(generator=g)
    t1, t2 = spec.cfg_scalars(cfg, [si], K)
    got, _, _ = spec.sample(lg, [0, L], t1, t2, tk, tp, noise, want_mixed=False)
    return ci, int((got != ref).sum()), B * L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-config", type=int, default=262144)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "flip_rate_r02.json"))
    a = ap.parse_args()
    B = 16
    per = B * L
    n_chunks = (a.rows_per_config + per - 1) // per
    jobs = [(ci, s, B) for ci in range(len(CONFIGS)) for s in range(n_chunks)]
    t0 = time.time()
    flips, rows = [0] * len(CONFIGS), [0] * len(CONFIGS)
    with mp.get_context("spawn").Pool(a.procs) as pool:
        for ci, f, n in pool.imap_unordered(chunk, jobs, chunksize=4):
            flips[ci] += f
            rows[ci] += n
    out = {"what": "tokens of oracle/spec_c sdvar_spec_sample vs reference models/helpers.py:6-19 on identical logits + noise",
           "V": V, "cfg": 1.5, "stage": "si=6 of K=10", "torch": torch.__version__, "seconds": round(time.time() - t0, 1),
           "configs": [{"logit_scale": c[0], "top_k": c[1], "top_p": c[2], "rows": rows[i], "flips": flips[i],
                        "flip_rate": flips[i] / max(rows[i], 1)} for i, c in enumerate(CONFIGS)],
           "total_rows": sum(rows), "total_flips": sum(flips)}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
