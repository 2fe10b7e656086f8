"""GPU parity of the whole path through the public Python API (VAR / SDVAR), against the oracle.

Token identity is only meaningful on identical input logits: bf16 GEMMs move logits by ~1e-3
relative, which flips argmax(p/noise) for a fraction of tokens and then changes every later stage.  So end-to-end
tests are TEACHER-FORCED: the oracle is run on the engine's own tokens, logits are compared with a stated
tolerance, and token / accept decisions are compared bit-exactly per kernel on the engine's own logits.
"""
import numpy as np
import pytest
import torch

from sdvar_b200.weights import hashed, var_state_dict, vqvae_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda"
P4 = (1, 2, 3, 4)
P256 = (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)


def _build(pns, depth_d, depth_t, ch=32, shared_t=False, sd_device="cpu", **kw):
    from sdvar_b200.models import build_vae_var_speculative_decoding
    vae, d, t, sd = build_vae_var_speculative_decoding(DEV, patch_nums=pns, ch=ch, depth_draft=depth_d, depth_target=depth_t,
                                                       shared_aln_target=shared_t)
    sds = dict(vae=vqvae_state_dict(ch=ch, patch_nums=pns, device=sd_device),
               d=var_state_dict(depth_d, patch_nums=pns, seed=1, tag="draft", device=sd_device, **kw),
               t=var_state_dict(depth_t, patch_nums=pns, seed=2, tag="target", shared_aln=shared_t, device=sd_device, **kw))
    vae.load_state_dict(sds["vae"]); d.load_state_dict(sds["d"]); t.load_state_dict(sds["t"])
    return vae, d, t, sd, sds


def _teacher_input(vq, idxs):
    """oracle restatement of idxBl_to_var_input (models/quant.py:169-184)"""
    B = idxs[0].shape[0]
    HW = vq.patch_nums[-1]
    f = torch.zeros(B, vq.Cvae, HW, HW)
    out = []
    for si in range(len(vq.patch_nums) - 1):
        f, nm = vq.next_input(si, f, idxs[si])
        out.append(nm.reshape(B, vq.Cvae, -1).transpose(1, 2))
    return torch.cat(out, 1), f


@pytest.mark.parametrize("pns,depth,shared", [(P4, 2, False), (P256, 4, False), (P256, 3, True)])
def test_teacher_forced_logits_vs_oracle(cuda_lib, pns, depth, shared):
    """VAR.forward (one block-causal window over all stages) against the oracle in both arithmetic regimes."""
    from oracle.ref_model import RefVAR
    vae, _, t, _, sds = _build(pns, 2, depth, shared_t=shared, gamma_bias=0.5, init_head=1.0)
    B = 2
    lab = torch.tensor([7, 421])
    L = sum(p * p for p in pns)
    x_in = torch.randn(B, L - 1, 32, generator=torch.Generator().manual_seed(0))
    got = t(lab.to(DEV), x_in.to(DEV)).cpu()
    ref16 = RefVAR(sds["t"], pns, mm="bf16").forward_teacher(lab, x_in)
    ref32 = RefVAR(sds["t"], pns, mm="fp32").forward_teacher(lab, x_in)
    scale = float(ref32.abs().max())
    # vs the bf16-emulating oracle: same rounding points, but accumulation order / exp differ by fp32 ulps, which flips
    # individual bf16 roundings (2^-8 relative each) of activations; through the blocks that is ~1e-3 of the logit range
    err = (got - ref16).abs()
    assert float(err.max()) < 8e-3 * scale and float(err.mean()) < 1e-3 * scale, (float(err.max()), float(err.mean()), scale)
    # vs the fp32 reference regime: bf16 operand rounding, a few 1e-2 of the logit range (SURVEY.md A1 anchor: 0.5%)
    assert float((got - ref32).abs().max()) < 3e-2 * scale, (float((got - ref32).abs().max()), scale)


def test_incremental_equals_window_pass_bitwise(cuda_lib):
    """pin P3 on the device: KV-cached stage-by-stage logits == one block-causal pass over all stages, bit for bit
    (every kernel is row-independent, so the window composition must not change a single row)."""
    vae, _, t, _, sds = _build(P256, 2, 3, gamma_bias=0.5, init_head=1.0)
    B = 2
    lab = torch.tensor([1, 2], device=DEV)
    L = t.L
    x_in = torch.randn(B, L - 1, 32, generator=torch.Generator().manual_seed(1)).to(DEV)
    full = t(lab, x_in)
    e = t._engine
    e.begin(B, lab)
    for si in range(len(P256)):
        l = t.ls[si]
        if si == 0:
            e.put_first_map(l)
        else:
            e.put_embed_map(si, x_in[:, t.begins[si] - 1:t.ends[si] - 1].transpose(1, 2).contiguous(), l)
        lg = e.forward([si])[:B]
        assert torch.equal(lg, full[:, t.begins[si]:t.ends[si]]), si


@pytest.mark.parametrize("top_k,top_p", [(0, 0.0), (900, 0.96)])
def test_autoregressive_infer_cfg_teacher_forced_parity(cuda_lib, top_k, top_p):
    from oracle import spec
    from oracle.ref_model import RefVAR, RefVQ, RefDecoder, ReplayNoise, cfg_mix
    vae, _, t, _, sds = _build(P256, 2, 3, gamma_bias=0.5, init_head=1.0)
    B, cfg = 3, 1.5
    lab = torch.tensor([3, 977, 500])
    rec = {}
    img, idxs, f_hat = t.autoregressive_infer_cfg(B, lab.to(DEV), cfg=cfg, top_k=top_k, top_p=top_p,
                                                  noise=ReplayNoise(5, DEV), return_tokens=True, record=rec)
    idxs = [i.cpu() for i in idxs]
    K = len(P256)
    # (1) K3 on the engine's own logits: bit-exact tokens vs the C spec
    for si in range(K):
        t1, t2 = spec.cfg_scalars(cfg, [si], K)
        ref_idx, _, _ = spec.sample(rec["logits"][si].cpu(), [0, t.ls[si]], t1, t2, top_k, top_p, rec["noise"][si].cpu(), want_mixed=False)
        assert torch.equal(ref_idx, idxs[si]), si
    # (2) f_hat / next maps: oracle VQ (the reference's own F.interpolate ops) on the engine's tokens
    vq = RefVQ(sds["vae"], P256)
    x_in, f_ref = _teacher_input(vq, idxs)
    f_ref, _ = vq.next_input(K - 1, f_ref, idxs[-1])
    assert torch.allclose(f_hat.cpu(), f_ref, rtol=1e-4, atol=1e-4)
    # (3) logits: oracle teacher-forced on the engine's tokens (bf16-emulating regime)
    o = RefVAR(sds["t"], P256, mm="bf16")
    both = torch.cat((lab, torch.full_like(lab, 1000)))
    ref_logits = o.forward_teacher(both, x_in.repeat(2, 1, 1))
    got = torch.cat([l.cpu() for l in rec["logits"]], dim=1)
    scale = float(ref_logits.abs().max())
    err = (got - ref_logits).abs()
    assert float(err.max()) < 8e-3 * scale and float(err.mean()) < 1e-3 * scale, (float(err.max()), float(err.mean()), scale)
    # (4) image: bf16 channels-last decoder vs the fp32 oracle decoder on the same f_hat; images live in [0,1]
    ref_img = RefDecoder(sds["vae"]).fhat_to_img(f_hat.cpu()).add_(1).mul_(0.5)
    assert img.shape == (B, 3, 256, 256)
    assert float((img.cpu() - ref_img).abs().mean()) < 1e-2 and float((img.cpu() - ref_img).abs().max()) < 0.15


def test_sd_test3_identities(cuda_lib):
    """pins P1/P2 (models/var.py:605-865): entry_num=0 == target baseline, entry_num=K == draft baseline, bit-exact,
    when fed the same noise."""
    from oracle.ref_model import ReplayNoise
    vae, d, t, sd, _ = _build(P4, 2, 3, gamma_bias=0.5, init_head=1.0)
    B, lab = 2, torch.tensor([5, 6], device=DEV)
    kw = dict(cfg=1.5, top_k=900, top_p=0.96, return_tokens=True)
    _, i0, f0 = sd.sdvar_autoregressive_infer_cfg_sd_test3(B, lab, entry_num=0, noise=ReplayNoise(1, DEV), **kw)
    _, it, ft = t.autoregressive_infer_cfg(B, lab, noise=ReplayNoise(1, DEV), **kw)
    assert all(torch.equal(a, b) for a, b in zip(i0, it)) and torch.equal(f0, ft)
    _, iK, fK = sd.sdvar_autoregressive_infer_cfg_sd_test3(B, lab, entry_num=len(P4), noise=ReplayNoise(1, DEV), **kw)
    _, idr, fd = d.autoregressive_infer_cfg(B, lab, noise=ReplayNoise(1, DEV), **kw)
    assert all(torch.equal(a, b) for a, b in zip(iK, idr)) and torch.equal(fK, fd)
    # hand-over in the middle: target continues from the draft's f_hat with an empty cache
    img, im, fm = sd.sdvar_autoregressive_infer_cfg_sd_test3(B, lab, entry_num=2, noise=ReplayNoise(1, DEV), **kw)
    assert all(torch.equal(a, b) for a, b in zip(im[:2], idr[:2])) and img.shape == (B, 3, 64, 64)
    with pytest.raises(NotImplementedError):
        sd.sdvar_autoregressive_infer_cfg_sd_test3(B, lab, entry_num=2, sd_mask=3)


class _RenameNoise:
    """route the SD loop's 'draft' stream to the baseline's 'target' stream so both consume the same tensors"""
    def __init__(self, inner):
        self.inner = inner
    def exponential(self, stream, rows, V):
        return self.inner.exponential("draft" if stream == "target" else stream, rows, V)
    def uniform(self, stream, rows):
        return self.inner.uniform(stream, rows)


@pytest.mark.parametrize("gamma", [1, 2, 3, 10])
def test_sd_loop_draft_equals_target_accepts_everything(cuda_lib, gamma):
    """draft == target => p == q bit for bit (row-independent kernels), so every token is accepted, every round commits
    gamma stages, and the result equals the baseline loop fed the draft-stream noise."""
    from oracle.ref_model import ReplayNoise
    from sdvar_b200.models import SDVAR
    vae, d, t, _, _ = _build(P256, 3, 3, gamma_bias=0.5, init_head=1.0)
    t.load_state_dict(d.state_dict())
    sd = SDVAR(d, t)
    B, lab = 2, torch.tensor([11, 12], device=DEV)
    kw = dict(cfg=1.5, top_k=900, top_p=0.96, return_tokens=True)
    img, idx_sd, f_sd = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, gamma=gamma, noise=ReplayNoise(3, DEV), **kw)
    st = sd.last_stats
    K = len(P256)
    assert st["rejected_tokens"] == 0 and st["accepted_tokens"] == B * t.L
    assert st["rounds"] == -(-K // gamma) and sum(st["advance"]) == K and st["target_passes"] == st["rounds"]
    _, idx_b, f_b = d.autoregressive_infer_cfg(B, lab, noise=_RenameNoise(ReplayNoise(3, DEV)), **kw)
    assert all(torch.equal(a, b) for a, b in zip(idx_sd, idx_b)) and torch.equal(f_sd, f_b)


@pytest.mark.parametrize("pair,gamma", [("equal", 3), ("close", 2), ("close", 3), ("far", 2), ("far", 4)])
def test_lazy_verify_equals_window_verify(cuda_lib, pair, gamma):
    """verify_mode='lazy' / 'auto' (stage-by-stage target verification with early exit) must be a pure scheduling change:
    tokens, f_hat, image, advances and acceptance counters identical to the one-pass window verification on the same noise,
    for a draft that equals the target (every stage accepted: lazy runs all g stages), one that is close to it (mixed rounds)
    and one that is far (every round repairs its first stage: lazy skips the rest of each window)."""
    from oracle.ref_model import ReplayNoise
    from sdvar_b200.models import SDVAR
    vae, d, t, sd, _ = _build(P256, 3, 3, gamma_bias=0.5, init_head=1.0)
    if pair != "far":
        sdict = {k: v.clone() for k, v in d.state_dict().items()}
        if pair == "close":
            sdict["head.bias"] = sdict["head.bias"] + 0.05 * hashed("lazy.hb", 0, tuple(sdict["head.bias"].shape), 1.0).to(DEV)
        t.load_state_dict(sdict)
    sd = SDVAR(d, t)
    B, lab = 3, torch.tensor([5, 6, 7], device=DEV)
    kw = dict(gamma=gamma, cfg=1.5, top_k=900, top_p=0.96, return_tokens=True)
    res = {}
    for mode in ("window", "lazy", "auto"):
        sd.LAZY_MIN_ROWS = 2 * B * 16                    # 'auto': stages of 16+ tokens verify lazily, smaller ones by window
        img, idxs, f_hat = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, noise=ReplayNoise(21, DEV), verify_mode=mode, **kw)
        res[mode] = (img.clone(), [i.clone() for i in idxs], f_hat.clone(), dict(sd.last_stats))
    w = res["window"]
    K = len(P256)
    for mode in ("lazy", "auto"):
        r = res[mode]
        assert all(torch.equal(a, b) for a, b in zip(r[1], w[1])), mode
        assert torch.equal(r[2], w[2]) and torch.equal(r[0], w[0]), mode
        for k in ("advance", "accepted_tokens", "rejected_tokens", "rounds", "stage_accept_tokens", "stage_tokens"):
            assert r[3][k] == w[3][k], (mode, k, r[3][k], w[3][k])
        assert r[3]["verify_mode"] == mode and sum(r[3]["advance"]) == K
    lz, wn = res["lazy"][3], w[3]
    windows = [min(gamma, K - s) for s in np.cumsum([0] + wn["advance"][:-1])]
    assert wn["target_passes"] == wn["rounds"] and wn["target_stages_skipped"] == 0
    # lazy: one single-stage pass per verified stage = the committed stages, except that a window accepted whole needs no repair pass
    assert lz["target_passes"] + lz["target_stages_skipped"] == sum(windows)
    assert lz["draft_stages"] == lz["target_passes"] and wn["draft_stages"] == sum(windows)   # lazy drafts a stage only when it verifies it
    if pair == "equal":
        assert lz["target_stages_skipped"] == 0 and lz["rejected_tokens"] == 0
    if pair == "far":
        assert lz["target_passes"] == K and lz["target_stages_skipped"] == sum(windows) - K and lz["rejected_tokens"] > 0
    if pair == "close":
        assert 0 < lz["rejected_tokens"] and max(lz["advance"]) > 1, "the close pair should mix accepted and repaired stages"


@pytest.mark.parametrize("rule,gamma", [("speculative", 2), ("speculative", 3), ("reference", 2)])
def test_sd_loop_invariants_after_rejections(cuda_lib, rule, gamma):
    """draft != target: rejections, repairs and KV rollback happen.  Invariants that pin the state handling:
    (a) f_hat == VQ(final tokens) (no double add, D7); (b) after the loop the target's KV cache equals, bit for bit, the
    cache of one clean teacher-forced pass over the final tokens (rollback leaves no stale or missing rows, D4);
    (c) bookkeeping: advances sum to K, one target pass per round."""
    from oracle.ref_model import RefVQ, ReplayNoise
    vae, d, t, sd, sds = _build(P256, 2, 3, gamma_bias=0.5, init_head=1.0)
    B, lab = 3, torch.tensor([1, 2, 3], device=DEV)
    img, idxs, f_hat = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, gamma=gamma, cfg=1.5, top_k=0, top_p=0.0,
                                                                     accept_rule=rule, noise=ReplayNoise(9, DEV), return_tokens=True)
    st = sd.last_stats
    K = len(P256)
    assert len(idxs) == K and [i.shape[1] for i in idxs] == t.ls
    assert sum(st["advance"]) == K and st["target_passes"] == st["rounds"] == len(st["advance"])
    if rule == "speculative":
        assert st["rejected_tokens"] > 0, "random-init d2 vs d3 must produce rejections"
    assert torch.isfinite(img).all() and float(img.min()) >= 0 and float(img.max()) <= 1
    vq = RefVQ(sds["vae"], P256)
    x_in, f_ref = _teacher_input(vq, [i.cpu() for i in idxs])
    f_ref, _ = vq.next_input(K - 1, f_ref, idxs[-1].cpu())
    assert torch.allclose(f_hat.cpu(), f_ref, rtol=1e-4, atol=1e-4)
    e = t._engine
    assert e.kv_len == t.L
    k_loop = [k.clone() for k in e.k_cache]
    v_loop = [v.clone() for v in e.v_cache]
    # clean pass: same stage inputs rebuilt by the device VQ path from the final tokens
    vqd = vae.quantize
    fh = torch.zeros(B, 32, 16, 16, device=DEV)
    e.begin(B, lab)
    nm = None
    for si in range(K):
        l = t.ls[si]
        e.put_first_map(l) if si == 0 else e.put_embed_map(si, nm, l)
        e.forward([si], want_logits=False)
        fh, nm = vqd.next_input_from_idx(si, fh, idxs[si])
    for i in range(t.depth):
        assert torch.equal(k_loop[i][:, :, :t.L], e.k_cache[i][:, :, :t.L]), i
        assert torch.equal(v_loop[i][:, :, :, :t.L], e.v_cache[i][:, :, :, :t.L]), i


def test_512px_pyramid_shared_aln_target_sd_loop(cuda_lib):
    """BASELINE.json configs[3] shape at toy depth: 512 px pyramid (patch_nums up to 32, L=2240), target with shared adaLN
    (the d36 layout), draft without (fixes D10/D11).  Invariants as above + teacher-forced logits vs the oracle."""
    from oracle.ref_model import RefVAR, RefVQ, ReplayNoise
    P512 = (1, 2, 3, 4, 6, 9, 13, 18, 24, 32)
    vae, d, t, sd, sds = _build(P512, 2, 2, shared_t=True, gamma_bias=0.5, init_head=1.0)
    B, lab = 2, torch.tensor([10, 20], device=DEV)
    img, idxs, f_hat = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, gamma=2, cfg=1.5, top_k=900, top_p=0.96,
                                                                     noise=ReplayNoise(4, DEV), return_tokens=True)
    K = len(P512)
    assert img.shape == (B, 3, 512, 512) and bool(torch.isfinite(img).all())
    assert sum(sd.last_stats["advance"]) == K and [i.shape[1] for i in idxs] == t.ls and t.L == 2240
    vq = RefVQ(sds["vae"], P512)
    x_in, f_ref = _teacher_input(vq, [i.cpu() for i in idxs])
    f_ref, _ = vq.next_input(K - 1, f_ref, idxs[-1].cpu())
    assert torch.allclose(f_hat.cpu(), f_ref, rtol=1e-4, atol=1e-4)
    got = t(lab, x_in.to(DEV)).cpu()
    ref = RefVAR(sds["t"], P512, mm="bf16").forward_teacher(lab.cpu(), x_in)
    scale = float(ref.abs().max())
    err = (got - ref).abs()
    assert float(err.max()) < 8e-3 * scale and float(err.mean()) < 1e-3 * scale, (float(err.max()), float(err.mean()), scale)


@pytest.mark.parametrize("name,depth,C,H,seed", [("d16", 16, 1024, 16, 1), ("w30", 2, 1920, 30, 2)])
def test_real_width_logits_vs_reference_golden(cuda_lib, name, depth, C, H, seed):
    """north-star WIDTHS on the device: VAR.forward of VAR-d16 (C=1024, H=16, 16 blocks) and of a 2-block model at the d30
    width (C=1920, H=30) against (a) the REAL reference's fp32 logits (tests/golden/real_width.npz) within the bf16-vs-fp32
    tolerance of SURVEY.md A1 and (b) the bf16-emulating oracle within the tight tolerance that catches indexing bugs."""
    import os
    from oracle.ref_model import RefVAR
    from sdvar_b200.models.var import VAR
    from sdvar_b200.models.vqvae import VQVAE
    from sdvar_b200.weights import hashed
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "real_width.npz"))
    vae = VQVAE(vocab_size=4096, z_channels=32, ch=32, v_patch_nums=P256).to(DEV)
    m = VAR(vae_local=vae, depth=depth, embed_dim=C, num_heads=H, attn_l2_norm=True, patch_nums=P256).to(DEV)
    sd = var_state_dict(depth, patch_nums=P256, seed=seed, tag=name, embed_dim=C, num_heads=H, gamma_bias=0.5, init_head=1.0)
    m.load_state_dict(sd)
    lab = torch.tensor([int(z[f"{name}_label"])])
    x_in = hashed(f"golden.width.{name}.x", 0, (1, 679, 32), 1.0)
    got = m(lab.to(DEV), x_in.to(DEV)).cpu()
    scale = float(z[f"{name}_absmax"])
    ref32 = torch.from_numpy(z[f"{name}_logits_slice"])
    e32 = (got[:, :, :96] - ref32).abs()
    assert float(e32.max()) < 3e-2 * scale and float(e32.mean()) < 4e-3 * scale, (float(e32.max()), float(e32.mean()), scale)
    ref16 = RefVAR(sd, P256, num_heads=H, mm="bf16").forward_teacher(lab, x_in)
    e16 = (got - ref16).abs()
    assert float(e16.max()) < 8e-3 * scale and float(e16.mean()) < 1e-3 * scale, (float(e16.max()), float(e16.mean()), scale)


def _dense_aux(rec, n):
    """stage-major u / noise of a recorded round -> dense (b, pos) order for the C spec"""
    seg = rec["seg"]
    Lw, V = seg[-1], rec["noise"].shape[1]
    u, nz = torch.empty(n, Lw), torch.empty(n, Lw, V)
    for j in range(len(seg) - 1):
        l = seg[j + 1] - seg[j]
        u[:, seg[j]:seg[j + 1]] = rec["u"][n * seg[j]:n * seg[j + 1]].cpu().view(n, l)
        nz[:, seg[j]:seg[j + 1]] = rec["noise"][n * seg[j]:n * seg[j + 1]].cpu().view(n, l, V)
    return u, nz.view(n * Lw, V)


@pytest.mark.parametrize("schedule,gamma,top_k,top_p,depths", [("lockstep", 2, 900, 0.96, (2, 3)), ("lockstep", 3, 0, 0.0, (2, 3)),
                                                                  ("ragged", 2, 0, 0.0, (2, 3)), ("lockstep", 2, 900, 0.96, (16, 20)),
                                                                  ("ragged", 2, 0, 0.0, (16, 20))])
def test_sd_loop_replay_against_spec(cuda_lib, schedule, gamma, top_k, top_p, depths):
    """LOOP-LEVEL replay (parity gap 4): every round's verify inputs (mixed target / draft logits, draft tokens, u,
    resample noise) are recorded from the device loop and pushed through the C spec and the loop spec's advance rule
    (DESIGN.md 3.4): accept flags, repaired tokens, first-reject scan and accepted-prefix lengths must be identical, the
    committed tokens must be the draft tokens of the intact stages + the spec's output for the last committed stage, and
    the stage pointers must advance by min(#intact leading stages + 1, g) -- per batch (lock-step) or per image (ragged).
    depths (16, 20) is BASELINE.json configs[0] as stated: VAR-d16 draft + VAR-d20 target, 256 px, batch 4, cfg 1.5 (default init)."""
    from oracle import spec
    from oracle.ref_model import ReplayNoise
    kw = dict(gamma_bias=0.5, init_head=1.0) if depths == (2, 3) else {}
    vae, d, t, sd, _ = _build(P256, depths[0], depths[1], sd_device="cpu" if depths == (2, 3) else DEV, **kw)
    B, lab = 4, torch.tensor([1, 2, 3, 4], device=DEV)
    rec = {}
    _, final, _ = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, gamma=gamma, cfg=1.5, top_k=top_k, top_p=top_p, schedule=schedule,
                                                               noise=ReplayNoise(21, DEV), return_tokens=True, record=rec)
    K = len(P256)
    stage = [0] * B
    assert len(rec["rounds"]) == sd.last_stats["target_passes"]
    for r in rec["rounds"]:
        ids = list(range(B)) if r["group"] is None else r["group"].tolist()
        n, seg, s = len(ids), r["seg"], r["stage"]
        g = len(seg) - 1
        assert all(stage[b] == s for b in ids), "a group must hold images that share a stage"
        u, nz = _dense_aux(r, n)
        ref = spec.verify(r["xt"].cpu(), r["xd"].cpu(), r["d"].cpu(), u, nz, seg)
        for k, name in (("out_idx", "out"), ("accept", "accept"), ("first_reject", "first_reject"), ("n_accept", "n_accept"),
                        ("accepted_stages", "accepted_stages"), ("summary", "summary")):
            assert torch.equal(r[name].cpu(), ref[k]), (k, s)
        ok = ref["accepted_stages"].tolist()
        adv = [min(min(ok) + 1, g)] * n if schedule == "lockstep" else [min(a + 1, g) for a in ok]
        for i, b in enumerate(ids):
            a = adv[i]
            for j in range(a - 1):       # intact stages keep the draft tokens
                assert torch.equal(final[s + j][b].cpu(), r["d"][i, seg[j]:seg[j + 1]].cpu()), (s, j, b)
            assert torch.equal(final[s + a - 1][b].cpu(), ref["out_idx"][i, seg[a - 1]:seg[a]]), (s, b)
            stage[b] += a
    assert stage == [K] * B


@pytest.mark.parametrize("gamma", [2, 3])
def test_ragged_schedule_half_batch_identical_models(cuda_lib, gamma):
    """Per-image ragged acceptance (VERDICT r1 'done means').  The target is a copy of the draft whose class
    embedding differs for labels >= 500 only: images with labels < 500 see p == q bit for bit (every token accepted, they
    advance gamma stages per round), the others see different distributions and advance by their own accepted prefix.
    Checked: the identical-model images finish in ceil(K/gamma) rounds with zero rejected tokens while the batch as a whole
    needs more rounds; f_hat == VQ(final tokens); the target's KV cache equals, bit for bit and per image, a clean
    teacher-forced pass over the final tokens; the lock-step schedule on the same inputs needs at least as many target passes
    per image."""
    from oracle.ref_model import RefVQ, ReplayNoise
    from sdvar_b200.models import SDVAR
    vae, d, t, _, sds = _build(P256, 3, 3, gamma_bias=0.5, init_head=1.0)
    tsd = {k: v.clone() for k, v in d.state_dict().items()}
    tsd["class_emb.weight"][500:1000] += 0.5 * torch.randn(500, tsd["class_emb.weight"].shape[1], generator=torch.Generator().manual_seed(0)).to(DEV)
    t.load_state_dict(tsd)
    sd = SDVAR(d, t)
    B = 6
    lab = torch.tensor([3, 700, 41, 900, 77, 650], device=DEV)
    same = [0, 2, 4]
    rec = {}
    img, final, f_hat = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, gamma=gamma, cfg=1.5, top_k=0, top_p=0.0, schedule="ragged",
                                                                     noise=ReplayNoise(5, DEV), return_tokens=True, record=rec)
    st = dict(sd.last_stats)
    K = len(P256)
    # per-image history from the record
    rounds_of = {b: 0 for b in range(B)}
    rejected_of = {b: 0 for b in range(B)}
    for r in rec["rounds"]:
        ids = list(range(B)) if r["group"] is None else r["group"].tolist()
        for i, b in enumerate(ids):
            rounds_of[b] += 1
            ok = int(r["accepted_stages"][i])
            g = len(r["seg"]) - 1
            upto = r["seg"][min(ok + 1, g)]
            rejected_of[b] += int((r["accept"][i, :upto] == 0).sum())
    for b in same:
        assert rounds_of[b] == -(-K // gamma) and rejected_of[b] == 0, (b, rounds_of[b], rejected_of[b])
    assert max(rounds_of.values()) > -(-K // gamma), "the perturbed images must have rejected something"
    assert st["rounds"] == max(rounds_of.values())
    assert torch.isfinite(img).all()
    vq = RefVQ(sds["vae"], P256)
    x_in, f_ref = _teacher_input(vq, [i.cpu() for i in final])
    f_ref, _ = vq.next_input(K - 1, f_ref, final[-1].cpu())
    assert torch.allclose(f_hat.cpu(), f_ref, rtol=1e-4, atol=1e-4)
    e = t._engine
    k_loop = [k.clone() for k in e.k_cache]
    v_loop = [v.clone() for v in e.v_cache]
    vqd = vae.quantize
    fh = torch.zeros(B, 32, 16, 16, device=DEV)
    e.begin(B, lab)
    nm = None
    for si in range(K):
        l = t.ls[si]
        e.put_first_map(l) if si == 0 else e.put_embed_map(si, nm, l)
        e.forward([si], want_logits=False)
        fh, nm = vqd.next_input_from_idx(si, fh, final[si])
    for i in range(t.depth):
        assert torch.equal(k_loop[i][:, :, :t.L], e.k_cache[i][:, :, :t.L]), i
        assert torch.equal(v_loop[i][:, :, :, :t.L], e.v_cache[i][:, :, :, :t.L]), i
    # lock-step on the same inputs: every image is held back by the slowest one
    sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, gamma=gamma, cfg=1.5, top_k=0, top_p=0.0, schedule="lockstep", noise=ReplayNoise(5, DEV))
    assert sd.last_stats["rounds"] >= max(rounds_of.values())


def test_gamma_policy_reference_shrinks_window(cuda_lib):
    """the reference's gamma controller (models/var.py:1352-1372): after a round in which no drafted stage survived intact
    the window shrinks by one, never below 1; with the default 'fixed' policy it stays put."""
    from oracle.ref_model import ReplayNoise
    vae, d, t, sd, _ = _build(P256, 2, 3, gamma_bias=0.5, init_head=1.0)
    B, lab = 3, torch.tensor([1, 2, 3], device=DEV)
    rec = {}
    sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, gamma=3, cfg=1.5, gamma_policy="reference", noise=ReplayNoise(9, DEV), record=rec)
    widths = [len(r["seg"]) - 1 for r in rec["rounds"]]
    g = 3
    for r, w in zip(rec["rounds"], widths):
        assert w == min(g, len(P256) - r["stage"])
        if int(r["summary"][0]) == 0:
            g = max(1, g - 1)
    assert widths[-1] <= widths[0] and min(widths) >= 1 and sum(sd.last_stats["advance"]) == len(P256)
    rec2 = {}
    sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, gamma=3, cfg=1.5, gamma_policy="fixed", noise=ReplayNoise(9, DEV), record=rec2)
    assert all(len(r["seg"]) - 1 == min(3, len(P256) - r["stage"]) for r in rec2["rounds"])


def test_decoder_per_layer_parity(cuda_lib):
    """fhat_to_img layer by layer (parity gap 5): the bf16 channels-last device decoder against the fp32 oracle
    (reference models/basic_vae.py:163-226) at conv_in, after the mid block, after every up level and at the image.  Each
    feature map must agree within bf16 round-off accumulated over the layers so far -- relative to that map's own scale --
    so an indexing bug in ONE layer cannot hide under the final image tolerance."""
    from oracle.ref_model import RefDecoder
    from sdvar_b200.models.vqvae import VQVAE
    sd = vqvae_state_dict(ch=32, patch_nums=P256)
    vae = VQVAE(vocab_size=4096, z_channels=32, ch=32, v_patch_nums=P256).to(DEV)
    vae.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
    f_hat = hashed("dec.fhat", 3, (2, 32, 16, 16), 1.0)
    ref_taps = {}
    ref_img = RefDecoder(sd).fhat_to_img(f_hat, taps=ref_taps)
    mods = vae._decoder_exec()
    dec = mods[1]
    got = {}
    hooks = [dec.conv_in.register_forward_hook(lambda m, i, o: got.__setitem__("conv_in_nobias", o)),
             dec.mid.register_forward_hook(lambda m, i, o: got.__setitem__("mid", o.float().cpu()))]
    for lv in range(len(dec.up)):
        hooks.append(dec.up[lv].register_forward_hook(lambda m, i, o, lv=lv: got.__setitem__(f"up.{lv}", o.float().cpu())))
    img = vae.fhat_to_img(f_hat.to(DEV)).cpu()
    for h in hooks:
        h.remove()
    assert set(ref_taps) - {"conv_in"} <= set(got)
    for k in ("mid", "up.4", "up.3", "up.2", "up.1", "up.0"):
        r, g = ref_taps[k], got[k]
        assert r.shape == g.shape, k
        scale = float(r.abs().max())
        err = (g - r).abs()
        assert float(err.max()) < 4e-2 * scale and float(err.mean()) < 4e-3 * scale, (k, float(err.max()), float(err.mean()), scale)
    err = (img - ref_img).abs()                      # images in [-1, 1]
    # 25 bf16 convolutions deep; the per-layer checks above carry the indexing guarantee, this one bounds the accumulated round-off
    assert float(err.max()) < 0.12 and float(err.mean()) < 1.2e-2, (float(err.max()), float(err.mean()))


def test_fid_pipeline_and_checkpoint_ingest(cuda_lib, tmp_path):
    """SURVEY.md 8f #4 at toy size: state dicts saved like the released checkpoints (torch.save of the reference's key
    surface) load with strict=True through sdvar_b200.fid.load_checkpoints; sample_fid_set batch-generates per-class images,
    packs them as the DiT-style npz (arr_0: (N,H,W,3) uint8, utils/misc.py:360-381) and writes PNGs that decode to the same
    pixels; the uint8 values are trunc(x*255) of the float images (sdvar_colab_test.py:235-236)."""
    from sdvar_b200 import fid
    pns = (1, 2, 3, 4)
    vae, d, t, sd, sds = _build(pns, 2, 3, gamma_bias=0.5, init_head=1.0)
    torch.save({k: v.cpu() for k, v in sds["vae"].items()}, tmp_path / "vae.pth")
    torch.save({k: v.cpu() for k, v in sds["d"].items()}, tmp_path / "var_d2.pth")
    torch.save({k: v.cpu() for k, v in sds["t"].items()}, tmp_path / "var_d3.pth")
    fid.load_checkpoints(vae, str(tmp_path / "vae.pth"), draft=(d, str(tmp_path / "var_d2.pth")), target=(t, str(tmp_path / "var_d3.pth")))
    kw = dict(cfg=1.5, top_k=900, top_p=0.96)
    floats = []

    def gen(B, lab, s):
        img = sd.sdvar_autoregressive_infer_cfg_parallel_v1(B, lab, g_seed=s, gamma=2, **kw)
        floats.append(img.clone())
        return img
    path = fid.sample_fid_set(gen, str(tmp_path / "s.npz"), classes=range(5), per_class=3, batch=4, device=DEV, png_dir=str(tmp_path / "png"))
    arr = np.load(path)["arr_0"]
    assert arr.shape == (15, 64, 64, 3) and arr.dtype == np.uint8
    ref = (torch.cat(floats).clamp(0, 1) * 255).to(torch.uint8).permute(0, 2, 3, 1).cpu().numpy()
    assert np.array_equal(arr, ref)
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(tmp_path / "png" / "000007.png")), arr[7])
