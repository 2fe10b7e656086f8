"""GPU tests at BASELINE.json's FULL sizes, through size-independent properties (the oracle cannot run these sizes in seconds):
the verify / sampling kernels on B x 680 x 4096 logits (configs[4]) and the whole draft-then-verify generation at B = 64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
P256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)
LS = [p * p for p in P256]
SEG = [0] + [int(v) for v in np.cumsum(LS)]


@pytest.mark.parametrize("B,scale", [(256, 0.05), (64, 3.0)])
def test_verify_full_size_properties(cuda_lib, B, scale):
    """K4 on B x 680 x 4096 (configs[4]): (1) the accept flag equals the accept test re-evaluated from the kernel's own p_d, q_d;
    (2) accepted rows keep the draft token, rejected rows return a token with residual mass p > q (recomputed in fp64);
    (3) per-(image, stage) accept counts / first-reject positions / accepted-prefix lengths / batch summary equal the same
    quantities recomputed from the flags with torch; (4) a second launch gives bit-identical outputs (determinism)."""
    L, V, S = 680, 4096, 10
    g = torch.Generator(device=DEV).manual_seed(5)
    xt = torch.randn(B, L, V, device=DEV, generator=g) * scale
    xd = xt + torch.randn(B, L, V, device=DEV, generator=g) * scale * 0.5
    d = torch.multinomial(xd.view(-1, V).softmax(-1), 1, generator=g).view(B, L)
    u = torch.rand(B, L, device=DEV, generator=g)
    noise = torch.empty(B * L, V, device=DEV).exponential_(generator=g)
    outs = []
    for _ in range(2):
        o = dict(idx=torch.empty(B, L, dtype=torch.int64, device=DEV), acc=torch.empty(B, L, dtype=torch.uint8, device=DEV),
                 p=torch.empty(B, L, device=DEV), q=torch.empty(B, L, device=DEV),
                 fr=torch.empty(B, S, dtype=torch.int32, device=DEV), na=torch.empty(B, S, dtype=torch.int32, device=DEV),
                 st=torch.empty(B, dtype=torch.int32, device=DEV), sm=torch.empty(4, dtype=torch.int32, device=DEV))
        ws = torch.zeros(cuda_lib.verify_workspace_ints(B, S), dtype=torch.int32, device=DEV)
        cuda_lib.verify_accept_resample(xt, xd, d, u, noise, B, L, V, SEG, o["idx"], o["acc"], o["p"], o["q"], o["fr"], o["na"], o["st"], o["sm"], ws)
        torch.cuda.synchronize()
        outs.append(o)
    a, b = outs
    for k in a:
        assert torch.equal(a[k], b[k]), k                                            # (4)
    acc = a["acc"].bool()
    assert torch.equal(acc, (u * a["q"]) < a["p"])                                    # (1) fp32 product, same rounding as the kernel
    assert torch.equal(a["idx"][acc], d[acc])                                         # (2a)
    rej = ~acc
    assert 0 < int(rej.sum()) < B * L
    rows = rej.view(-1).nonzero().view(-1)[:4096]                                     # (2b) on a sample of the rejected rows
    pt = xt.view(-1, V)[rows].double().softmax(-1)
    qd = xd.view(-1, V)[rows].double().softmax(-1)
    tok = a["idx"].view(-1)[rows]
    resid = (pt - qd).gather(1, tok.view(-1, 1)).view(-1)
    assert bool((resid > -1e-9).all()), float(resid.min())
    # draft probabilities the kernel reports agree with fp64 softmax to fp32 round-off
    assert torch.allclose(a["q"].view(-1)[rows].double(), qd.gather(1, d.view(-1)[rows].view(-1, 1)).view(-1), rtol=1e-5, atol=1e-9)
    for j in range(S):                                                                # (3)
        seg_acc = acc[:, SEG[j]:SEG[j + 1]]
        assert torch.equal(a["na"][:, j].long(), seg_acc.sum(1))
        first = torch.where(seg_acc.all(1), torch.full((B,), LS[j], device=DEV), (~seg_acc).float().argmax(1))
        assert torch.equal(a["fr"][:, j].long(), first)
    full = torch.stack([acc[:, SEG[j]:SEG[j + 1]].all(1) for j in range(S)], 1).long()
    prefix = full.cumprod(1).sum(1)
    assert torch.equal(a["st"].long(), prefix)
    assert a["sm"].tolist() == [int(prefix.min()), int(acc.sum()), int(rej.sum()), 0]


def test_sample_full_size_properties(cuda_lib):
    """K3 at B=64, one full stage of 256 tokens, top_k=900 / top_p=0.96: kept set is non-empty and has at most top_k entries,
    contains the row maximum, the removed low tail carries at most 1 - top_p of the top-k mass (fp64), the sampled token is a
    kept one and maximises p / noise over the kept set (fp64, relative 1e-5), and the launch is deterministic."""
    B, l, V, tk, tp = 64, 256, 4096, 900, 0.96
    g = torch.Generator(device=DEV).manual_seed(11)
    lg = torch.randn(2 * B, l, V, device=DEV, generator=g) * 2.0
    noise = torch.empty(B * l, V, device=DEV).exponential_(generator=g)
    thr = float(np.float32(1 - tp))
    res = []
    for _ in range(2):
        idx = torch.empty(B, l, dtype=torch.int64, device=DEV)
        mixed = torch.empty(B, l, V, device=DEV)
        cuda_lib.sample_cfg_topk_topp(lg, B, l, V, [0, l], [float(np.float32(2.5))], [float(np.float32(1.5))], tk, thr, noise, idx, mixed, None)
        torch.cuda.synchronize()
        res.append((idx, mixed))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    idx, mixed = res[0]
    raw = (np.float32(2.5) * lg[:B] - np.float32(1.5) * lg[B:]).view(-1, V)
    m = mixed.view(-1, V)
    kept = torch.isfinite(m)
    nk = kept.sum(1)
    assert int(nk.min()) >= 1 and int(nk.max()) <= tk
    assert torch.equal(m[kept], raw[kept])                                            # kept logits pass through unchanged
    assert bool(kept.gather(1, raw.argmax(1, keepdim=True)).all())                    # the maximum is always kept
    topk_mask = raw >= raw.topk(tk, dim=1).values[:, -1:]
    p_topk = torch.where(topk_mask, raw.double(), torch.full_like(raw, -float("inf"), dtype=torch.float64)).softmax(-1)
    removed_mass = (p_topk * (~kept & topk_mask)).sum(1)
    assert float(removed_mass.max()) <= (1 - tp) + 1e-6
    tok = idx.view(-1, 1)
    assert bool(kept.gather(1, tok).all())
    ratio = m.double().softmax(-1) / noise.double()
    assert bool((ratio.gather(1, tok).view(-1) >= ratio.max(1).values * (1 - 1e-5)).all())


def test_generation_full_batch_is_deterministic_and_consistent(cuda_lib):
    """The whole draft-then-verify loop at the BASELINE batch (B=64, 256 px pyramid, gamma=2, top_k=900/top_p=0.96) on a d16
    draft and a d20 target: two runs with the same seed give bit-identical tokens and images; the acceptance statistics are
    consistent (every committed token is either accepted or repaired, one target pass per round, stages advance to the end); images are finite and in [0, 1]."""
    from sdvar_b200.models import build_vae_var_speculative_decoding
    from sdvar_b200.weights import var_state_dict, vqvae_state_dict
    vae, draft, target, sd = build_vae_var_speculative_decoding(device=DEV, patch_nums=P256, depth_draft=16, depth_target=20, ch=32)
    vae.load_state_dict(vqvae_state_dict(ch=32, patch_nums=P256, device=DEV))
    draft.load_state_dict(var_state_dict(16, patch_nums=P256, seed=1, tag="draft", device=DEV))
    target.load_state_dict(var_state_dict(20, patch_nums=P256, seed=2, tag="target", device=DEV))
    labels = torch.randint(0, 1000, (64,), generator=torch.Generator().manual_seed(0)).to(DEV)
    runs = []
    for _ in range(2):
        img, toks, _ = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B=64, label_B=labels, g_seed=3, cfg=1.5, gamma=2, top_k=900, top_p=0.96,
                                                                     return_tokens=True)
        torch.cuda.synchronize()
        runs.append((img.clone(), [t.clone() for t in toks], dict(sd.last_stats)))
    (i0, t0, s0), (i1, t1, s1) = runs
    assert torch.equal(i0, i1) and all(torch.equal(a, b) for a, b in zip(t0, t1))
    assert s0 == s1
    assert [t.shape for t in t0] == [(64, l) for l in LS]
    assert all(int(t.min()) >= 0 and int(t.max()) < 4096 for t in t0)
    assert i0.shape == (64, 3, 256, 256) and bool(torch.isfinite(i0).all()) and 0.0 <= float(i0.min()) and float(i0.max()) <= 1.0
    assert sum(s0["advance"]) == len(LS) and s0["rounds"] == len(s0["advance"]) == s0["target_passes"]
    # every committed token is either an accepted draft token or a target repair, counted once
    assert s0["accepted_tokens"] + s0["rejected_tokens"] == 64 * sum(LS)
