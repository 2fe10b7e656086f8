"""CPU tests of the draft->verify loop SPEC (oracle.ref_model.sd_generate, PARITY UNPINNED by the reference) and of the
data-parallel plumbing (gloo, world_size 2)."""
import os

import numpy as np

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.ref_model import RefVAR, RefVQ, ReplayNoise, sd_generate
from sdvar_b200.weights import var_state_dict, vqvae_state_dict

P4 = (1, 2, 3, 4)
KW = dict(gamma_bias=0.5, init_head=1.0)


class _Rename:
    def __init__(self, inner): self.inner = inner
    def exponential(self, stream, rows, V): return self.inner.exponential("draft", rows, V)
    def uniform(self, stream, rows): return self.inner.uniform(stream, rows)


def _models():
    vq = RefVQ(vqvae_state_dict(ch=32, patch_nums=P4), P4)
    d = RefVAR(var_state_dict(2, patch_nums=P4, seed=1, tag="draft", **KW), P4)
    t = RefVAR(var_state_dict(3, patch_nums=P4, seed=2, tag="target", **KW), P4)
    return vq, d, t


@pytest.mark.parametrize("gamma", [1, 2, 4])
def test_draft_equal_target_accepts_everything_and_equals_baseline(gamma):
    vq, d, _ = _models()
    d2 = RefVAR(var_state_dict(2, patch_nums=P4, seed=1, tag="draft", **KW), P4)
    B, lab = 2, torch.tensor([4, 5])
    f, idxs, st = sd_generate(d, d2, vq, B, lab, ReplayNoise(2), cfg=1.5, gamma=gamma, top_k=900, top_p=0.96)
    # window pass vs incremental pass differ by fp32 re-association (1e-7): p/q may differ in the last ulp, which can
    # reject a token only if u*q lands within that ulp of p -- not with these seeds
    assert st["rejected_tokens"] == 0 and sum(st["advance"]) == 4 and st["rounds"] == -(-4 // gamma)
    noise = _Rename(ReplayNoise(2))
    nl = [noise.exponential("target", B * l, 4096) for l in d.ls]
    fb, ib = d.autoregressive_infer_cfg(vq, B, lab, cfg=1.5, top_k=900, top_p=0.96, noise=nl)
    assert all(torch.equal(a, b) for a, b in zip(idxs, ib)) and torch.allclose(f, fb)


@pytest.mark.parametrize("rule", ["speculative", "reference"])
def test_loop_invariants_with_rejections(rule):
    vq, d, t = _models()
    B, lab = 3, torch.tensor([1, 2, 3])
    f, idxs, st = sd_generate(d, t, vq, B, lab, ReplayNoise(7), cfg=1.5, gamma=2, accept_rule=rule)
    assert [i.shape for i in idxs] == [(B, l) for l in d.ls]
    assert sum(st["advance"]) == 4 and st["target_passes"] == st["rounds"] and all(1 <= a <= 2 for a in st["advance"])
    fr = torch.zeros(B, 32, 4, 4)
    for si, ix in enumerate(idxs):
        fr, _ = vq.next_input(si, fr, ix)
    assert torch.allclose(f, fr, atol=1e-6)                       # f_hat == VQ(final tokens): no double add (D7)
    assert d.kv_len() == 0 and t.kv_len() == 0                    # caches released


def test_window_pass_equals_incremental_pass():
    """pin P3 on the oracle: one block-causal pass over g stages == g KV-cached stage passes"""
    vq, _, t = _models()
    B, lab = 2, torch.tensor([9, 10])
    cond = t.cond(lab)
    xs = [t.first_map(cond)] + [t.embed_map(si, torch.randn(B, 32, P4[si], P4[si], generator=torch.Generator().manual_seed(si))) for si in (1, 2, 3)]
    t.kv_caching(True)
    inc = [t.get_logits(t.blocks(x, cond, None), cond) for x in xs]
    t.kv_caching(True)
    a = t.forward_window(0, xs[:2], cond)
    b = t.forward_window(2, xs[2:], cond)
    for got, want in zip(a + b, inc):
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    t.kv_caching(False)


# ---------------------------------------------------------------------------------------------- DP plumbing, gloo
def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdvar_b200 import parallel
    G = 5                                                        # ragged: 3 + 2
    labels = torch.arange(100, 100 + G)
    mine = parallel.shard_labels(labels)
    img = mine.float().view(-1, 1, 1, 1).expand(-1, 3, 2, 2).contiguous()
    allimg = parallel.gather_images(img, global_batch=G)
    stats = parallel.reduce_stats(dict(rounds=10, target_passes=10, draft_stages=19, accepted_tokens=100 * (rank + 1), rejected_tokens=rank), "cpu")
    even = parallel.gather_images(torch.full((2, 3, 2, 2), float(rank)))
    q.put((rank, mine.tolist(), allimg[:, 0, 0, 0].tolist(), stats, even[:, 0, 0, 0].tolist(), parallel.rank_seed(7)))
    dist.destroy_process_group()


def test_dp_sharding_gather_and_stats_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(30) for p in ps]
    assert res[0][1] == [100, 101, 102] and res[1][1] == [103, 104]
    for r in res:
        assert r[2] == [100.0, 101.0, 102.0, 103.0, 104.0]
        assert r[3] == dict(rounds=20, target_passes=20, draft_stages=38, accepted_tokens=300, rejected_tokens=1)
        assert r[4] == [0.0, 0.0, 1.0, 1.0]
    assert res[0][5] == 7 and res[1][5] == 8


def test_shard_range_partitions_exactly():
    from sdvar_b200.parallel import shard_range
    for G in (1, 7, 64, 512):
        for w in (1, 2, 4, 8):
            r = [shard_range(G, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == G and all(a[1] == b[0] for a, b in zip(r, r[1:]))


@pytest.mark.parametrize("policy", ["fixed", "reference"])
def test_ragged_schedule_spec(policy):
    """oracle of the per-image ragged schedule (oracle.ref_model.sd_generate_ragged): the target equals the draft except for
    the class embedding of labels >= 500, so images with smaller labels accept every drafted stage (ceil(K/gamma) rounds, the
    draft's own tokens) while the others advance by their own accepted prefixes; f_hat == VQ(final tokens) for every image; a
    B=1 lock-step run of an image that never rejects commits the same number of stages per round."""
    from oracle.ref_model import sd_generate_ragged
    vq, d, _ = _models()
    tsd = {k: v.clone() for k, v in var_state_dict(2, patch_nums=P4, seed=1, tag="draft", **KW).items()}
    tsd["class_emb.weight"][500:1000] += 0.5 * torch.randn(500, tsd["class_emb.weight"].shape[1], generator=torch.Generator().manual_seed(0))
    t = RefVAR(tsd, P4)
    B, lab = 4, torch.tensor([3, 700, 41, 900])
    f, idxs, st = sd_generate_ragged(d, t, vq, B, lab, ReplayNoise(3), cfg=1.5, gamma=2, gamma_policy=policy)
    K = len(P4)
    assert st["image_rounds"][0] == st["image_rounds"][2] == -(-K // 2)
    assert max(st["image_rounds"]) <= K and st["rounds"] == max(st["image_rounds"])
    fr = torch.zeros(B, 32, 4, 4)
    for si, ix in enumerate(idxs):
        fr, _ = vq.next_input(si, fr, ix)
    assert torch.allclose(f, fr, atol=1e-6)
    assert [i.shape for i in idxs] == [(B, l) for l in d.ls]


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdvar_b200 import parallel
    B = 3
    img = torch.rand(B, 3, 8, 8, generator=torch.Generator().manual_seed(rank))
    buf = parallel.GatherBuffer(world, B, (3, 8, 8), "cpu")
    out = buf.gather(img).clone()
    t = parallel.reduce_stats_async(dict(rounds=1 + rank, target_passes=2, draft_stages=3, accepted_tokens=10 * (rank + 1), rejected_tokens=1), "cpu")
    q.put((rank, out, parallel.stats_from_tensor(t)))
    dist.destroy_process_group()


def test_uint8_gather_and_async_stats_gloo_world2():
    """the path's only collectives on 2 gloo ranks: every rank's images land as uint8 = trunc(x*255) in its slot
    of one preallocated buffer, the acceptance counters are summed without a host synchronisation inside the step"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
    want = torch.cat([(torch.rand(3, 3, 8, 8, generator=torch.Generator().manual_seed(r)).clamp(0, 1) * 255).to(torch.uint8) for r in range(2)])
    for rank, out, st in res:
        assert out.dtype == torch.uint8 and torch.equal(out, want)
        assert st == dict(rounds=3, target_passes=4, draft_stages=6, accepted_tokens=30, rejected_tokens=2)


@pytest.mark.parametrize("pair,gamma", [("equal", 2), ("equal", 4), ("far", 2), ("far", 3)])
def test_lazy_schedule_equals_window_schedule(pair, gamma):
    """spec level (CPU, fp32): verify_mode='lazy' -- draft and verify stage by stage with early exit, the unreached stages' noise
    drawn and dropped -- commits the same tokens and advances as the one-pass window schedule.  (The oracle's window pass and
    its incremental passes agree to fp32 re-association, 1e-7, which moves a decision only on an exact near-tie -- not with
    these seeds; on the device the two passes are bit-identical and the same identity is asserted exactly in
    tests/test_engine_gpu.py::test_lazy_verify_equals_window_verify.)"""
    vq, d, t = _models()
    if pair == "equal":
        t = RefVAR(var_state_dict(2, patch_nums=P4, seed=1, tag="draft", **KW), P4)
    B, lab = 2, torch.tensor([4, 5])
    kw = dict(cfg=1.5, gamma=gamma, top_k=900, top_p=0.96)
    fw, iw, sw = sd_generate(d, t, vq, B, lab, ReplayNoise(5), verify_mode="window", **kw)
    fl, il, sl = sd_generate(d, t, vq, B, lab, ReplayNoise(5), verify_mode="lazy", **kw)
    assert all(torch.equal(a, b) for a, b in zip(iw, il)) and torch.allclose(fw, fl, atol=1e-6)
    for k in ("advance", "rounds", "accepted_tokens", "rejected_tokens", "stage_tokens", "stage_accept_tokens"):
        assert sw[k] == sl[k], k
    K = len(P4)
    windows = [min(gamma, K - s) for s in np.cumsum([0] + sw["advance"][:-1])]
    assert sw["draft_stages"] == sum(windows) and sw["target_passes"] == sw["rounds"]
    # lazy: one draft stage + one target pass per VERIFIED stage; every committed stage is verified, wasted ones are not
    assert sl["draft_stages"] == sl["target_passes"] and sl["target_passes"] <= sum(windows) and sl["target_passes"] >= K
    if pair == "far":
        assert sl["rejected_tokens"] > 0 and sl["target_passes"] < sum(windows)
