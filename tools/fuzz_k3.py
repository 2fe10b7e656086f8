#!/usr/bin/env python
"""Randomised bit-exactness sweep of K3 (sdvar_sample_cfg_topk_topp) against the C spec on one B200: random filter settings
(top_k in {0, 1..V-1}, top_p in {0, (0.01, 0.99999)}), CFG strengths and row shapes -- Gaussian at several scales, peaked,
heavy-tailed, quantised (tie groups around both cuts), duplicated plateaus, rows holding -inf.  Prints `bad 0` when every index,
probability and masked logit row is identical.  TEST INFRASTRUCTURE (imports oracle/).   python tools/fuzz_k3.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spec          # noqa: E402
from sdvar_b200 import _cabi     # noqa: E402

DEV = "cuda"


def rows(g, kind, n, V):
    x = torch.randn(n, V, generator=g)
    if kind == 0:
        return x * float(10 ** g_uniform(g, -2, 1))
    if kind == 1:      # peaked
        x = x * 2
        for r in range(n):
            x[r, torch.randint(0, V, (int(torch.randint(1, 6, (1,), generator=g)),), generator=g)] += float(g_uniform(g, 4, 30))
        return x
    if kind == 2:      # heavy tails
        return torch.sign(x) * x.abs() ** 3
    if kind == 3:      # quantised
        q = float(2 ** int(torch.randint(0, 5, (1,), generator=g)))
        return (x * 3 * q).round() / q
    if kind == 4:      # a plateau of equal values somewhere in the order statistics
        s = x.sort(dim=-1).values
        a = int(torch.randint(0, V - 600, (1,), generator=g)); w = int(torch.randint(2, 600, (1,), generator=g))
        return torch.where((x >= s[:, a:a + 1]) & (x <= s[:, a + w:a + w + 1]), s[:, a:a + 1], x)
    x[:, ::int(torch.randint(2, 9, (1,), generator=g))] = float("-inf")      # kind 5
    return x


def g_uniform(g, lo, hi):
    return lo + (hi - lo) * float(torch.rand(1, generator=g))


def run(n_cases, seed):
    g = torch.Generator().manual_seed(seed)
    bad = 0
    for case in range(n_cases):
        V = [4096, 4096, 4096, 1024, 2048, 8192][case % 6]
        B, L = 2, int(torch.randint(1, 9, (1,), generator=g))
        kind = case % 6 if case % 7 else 5
        cond = rows(g, kind, B * L, V).view(B, L, V)
        unc = rows(g, 0, B * L, V).view(B, L, V) if kind != 5 else torch.zeros(B, L, V)
        lg = torch.cat([cond, unc], 0).contiguous()
        top_k = 0 if case % 5 == 0 else int(torch.randint(1, V, (1,), generator=g))
        top_p = 0.0 if case % 4 == 1 else min(0.99999, max(0.01, g_uniform(g, 0.0, 1.05)))
        noise = torch.empty(B * L, V).exponential_(generator=g)
        t1, t2 = spec.cfg_scalars(g_uniform(g, 0.0, 4.0), [int(torch.randint(0, 10, (1,), generator=g))], 10)
        ri, rm, rp = spec.sample(lg, [0, L], t1, t2, top_k, top_p, noise)
        idx = torch.empty(B, L, dtype=torch.int64, device=DEV)
        mixed = torch.empty(B, L, V, device=DEV)
        prob = torch.empty(B, L, device=DEV)
        _cabi.sample_cfg_topk_topp(lg.to(DEV), B, L, V, [0, L], t1, t2, top_k, spec.top_p_threshold(top_p), noise.to(DEV), idx, mixed, prob)
        torch.cuda.synchronize()
        ok = torch.equal(mixed.cpu().view(torch.int32), rm.view(torch.int32)) and torch.equal(idx.cpu(), ri)
        ok = ok and torch.equal(prob.cpu().view(torch.int32), rp.view(torch.int32))
        if not ok:
            bad += 1
            print("FAIL case", case, "V", V, "kind", kind, "top_k", top_k, "top_p", top_p, "L", L,
                  "kept gpu/spec", int(torch.isfinite(mixed).sum()), int(torch.isfinite(rm).sum()),
                  "idx eq", bool(torch.equal(idx.cpu(), ri)))
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    print("cases", n, "bad", run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0))
