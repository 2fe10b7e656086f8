"""CPU tests of the plain-C arithmetic spec (oracle/spec_c): accuracy of the fixed exp, agreement of the verify
rule with its PyTorch restatement, scan semantics, edge cases."""
import numpy as np
import torch

from oracle import spec
from oracle.ref_model import filter_top_k_top_p_, first_reject_scan, verify_tokens
from sdvar_b200.weights import hashed


def test_expf_within_one_ulp_and_flushes():
    x = torch.linspace(-87.0, 0.0, 5001)
    e = spec.expf(x)
    ref = torch.exp(x.double())
    assert float(((e.double() - ref) / ref).abs().max()) < 1.2e-7
    assert float(spec.expf(torch.tensor([-200.0, float("-inf")])).abs().max()) == 0.0
    assert float(spec.expf(torch.tensor([0.0]))) == 1.0


def test_race_reciprocal_within_one_ulp():
    """sdvar_spec_rcp, the division-free reciprocal of the filtered regime's exponential race (magic seed + 4 Newton steps in
    fmaf): relative error below 2^-23 over Exp(1) draws and 80 binades -- the race compares e * R(noise) where the reference
    compares p / noise, and only a near-tie of that size can tell the two apart (pinned by the flip-rate test)"""
    import ctypes as C
    l = spec.lib()
    l.sdvar_spec_rcp.restype = C.c_float
    l.sdvar_spec_rcp.argtypes = [C.c_float]
    g = np.random.default_rng(0)
    xs = np.concatenate([g.exponential(size=20000).astype(np.float32),
                         (np.float32(2.0) ** g.integers(-40, 40, 2000) * g.uniform(1, 2, 2000)).astype(np.float32)])
    err = max(abs(float(l.sdvar_spec_rcp(float(x))) * float(x) - 1.0) for x in xs if x > 0)
    assert err < 2.0 ** -23


def _case(B, ls, V, scale, seed, top_k=0, top_p=0.0):
    L = sum(ls)
    g = torch.Generator().manual_seed(seed)
    xt = hashed(f"spec.t.{scale}", seed, (B, L, V), scale)
    xd = xt + hashed(f"spec.d.{scale}", seed, (B, L, V), scale * 0.35)
    filter_top_k_top_p_(xt, top_k, top_p); filter_top_k_top_p_(xd, top_k, top_p)
    d = torch.multinomial(xd.softmax(-1).view(-1, V), 1, generator=g).view(B, L)
    return xt, xd, d, torch.rand(B, L, generator=g), torch.empty(B * L, V).exponential_(generator=g)


def test_verify_matches_torch_restatement_and_scan():
    B, V, ls = 3, 4096, [1, 4, 9, 16]
    seg = [0] + list(np.cumsum(ls))
    for scale, tk, tp in ((3.0, 0, 0.0), (3.0, 900, 0.96), (0.05, 0, 0.0)):
        xt, xd, d, u, noise = _case(B, ls, V, scale, 11, tk, tp)
        out = spec.verify(xt, xd, d, u, noise, seg)
        o_ref, acc_ref, p_d, q_d = verify_tokens(xt.view(-1, V), xd.view(-1, V), d.view(-1), u.view(-1), noise)
        assert torch.equal(out["accept"].bool().view(-1), acc_ref)
        assert torch.equal(out["out_idx"].view(-1), o_ref)
        assert torch.allclose(out["p_d"].view(-1), p_d, rtol=2e-6) and torch.allclose(out["q_d"].view(-1), q_d, rtol=2e-6)
        acc = out["accept"].bool()
        for j in range(len(ls)):
            fr, na = first_reject_scan(acc[:, seg[j]:seg[j + 1]])
            assert torch.equal(out["first_reject"][:, j].long(), fr) and torch.equal(out["n_accept"][:, j].long(), na)
        lead = [next((j for j in range(len(ls)) if int(out["n_accept"][b, j]) != ls[j]), len(ls)) for b in range(B)]
        assert out["accepted_stages"].tolist() == lead
        assert out["summary"].tolist() == [min(lead), int(acc.sum()), int((~acc).sum()), 0]


def test_verify_accepts_everything_when_p_equals_q_and_rejects_out_of_support():
    B, V, ls = 2, 4096, [4, 9]
    seg = [0, 4, 13]
    xt, _, d, u, noise = _case(B, ls, V, 2.0, 3)
    out = spec.verify(xt, xt, d, u, noise, seg)
    assert bool(out["accept"].all()) and torch.equal(out["out_idx"], d) and out["summary"].tolist() == [2, 26, 0, 0]
    # a draft token the target filtered out (p[d] = 0) is always rejected and repaired inside the target's support
    xt2 = xt.clone()
    xt2.scatter_(-1, d.unsqueeze(-1), float("-inf"))
    out = spec.verify(xt2, xt, d, u, noise, seg)
    assert not bool(out["accept"].any())
    assert bool(torch.isfinite(xt2.gather(-1, out["out_idx"].unsqueeze(-1))).all())
    assert out["first_reject"].tolist() == [[0, 0], [0, 0]] and out["accepted_stages"].tolist() == [0, 0]


def test_sample_edges_topk_ge_V_and_single_survivor():
    B, L, V = 1, 3, 4096
    lg = hashed("spec.edge", 0, (2 * B, L, V), 1.0)
    noise = torch.empty(B * L, V).exponential_(generator=torch.Generator().manual_seed(0))
    t1, t2 = spec.cfg_scalars(0.0, [0], 10)
    idx_a, mixed_a, _ = spec.sample(lg, [0, L], t1, t2, 0, 0.0, noise)
    idx_b, mixed_b, _ = spec.sample(lg, [0, L], t1, t2, V, 0.0, noise)          # top_k == V keeps everything
    assert torch.equal(idx_a, idx_b) and torch.equal(mixed_a, mixed_b) and torch.equal(mixed_a, lg[:B])   # t=0: x = cond
    idx_c, mixed_c, prob_c = spec.sample(lg, [0, L], t1, t2, 1, 0.0, noise)      # top_k == 1: argmax survives alone
    assert torch.equal(idx_c, lg[:B].argmax(-1)) and int((~torch.isinf(mixed_c)).sum()) == L and torch.all(prob_c == 1.0)
    idx_d, mixed_d, _ = spec.sample(lg, [0, L], t1, t2, 0, 1e-6, noise)          # tiny top_p: only the max survives
    assert torch.equal(idx_d, lg[:B].argmax(-1))
