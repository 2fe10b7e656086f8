#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (/root/reference, read-only) -- run in the build
container only; the fixtures travel, the reference does not.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            # writes tests/golden/

All weights come from sdvar_b200.weights (bit-reproducible hashed init) loaded into the reference modules with
load_state_dict, so the fixtures hold only small inputs/outputs:
  sampler.npz    reference sample_with_top_k_top_p_ (models/helpers.py:6-19) on hashed logits: tokens + removed-entry masks
  vq.npz         reference get_next_autoregressive_input (models/quant.py:187-196) over a 256 px pyramid, B=1
  tiny_var.npz   reference VAR.autoregressive_infer_cfg / VAR.forward / SDVAR.sd_test3 on a depth-2/3 model, 4-stage pyramid
  blocks.npz     reference AdaLNSelfAttn stack (models/basic_var.py:152-159) input/output on the tiny model
  d16.npz        reference VAR-d16 autoregressive_infer_cfg, B=1, 256 px pyramid (BASELINE.json configs[0] size): tokens, f_hat
  encode.npz     reference encode side (`python oracle/make_golden.py encode` writes only this one): VectorQuantizer2.
                 f_to_idxBl_or_fhat (models/quant.py:135-166) on a hashed 16x16 feature map (tokens of all 10 scales, final
                 f_hat) and quant_conv(Encoder(img)) (models/vqvae.py:66, models/basic_vae.py:144-160) on a hashed image
  real_width.npz reference VAR.forward teacher-forced logits at the north-star widths (`... widths`): VAR-d16 and a 2-block
                 model at the d30 width (C=1920, H=30), active gates
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from sdvar_b200.weights import hashed, var_state_dict, vqvae_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
P4, P256 = (1, 2, 3, 4), (1, 2, 3, 4, 5, 6, 8, 10, 13, 16)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def build_ref(pns, depth, ch=32, shared_aln=False, **kw):
    with quiet():
        import models as R
        vae = R.VQVAE(vocab_size=4096, z_channels=32, ch=ch, test_mode=True, share_quant_resi=4, v_patch_nums=pns)
        var = R.VAR(vae_local=vae, depth=depth, embed_dim=64 * depth, num_heads=depth, attn_l2_norm=True, patch_nums=pns,
                    shared_aln=shared_aln, flash_if_available=False, fused_if_available=False)
    return vae, var.eval()


def load_vae(vae, vsd):
    r = vae.load_state_dict(vsd, strict=False)
    assert all(k.startswith(("encoder.", "quant_conv.")) for k in r.missing_keys) and not r.unexpected_keys


def encode_golden():
    """encode side"""
    os.makedirs(OUT, exist_ok=True)
    with quiet():
        import models as R
    o = {}
    # multi-scale residual quantisation of a feature map, 256 px pyramid, codebook / Phi from the hashed VQVAE recipe
    vsd = vqvae_state_dict(ch=32, patch_nums=P256, with_encoder=True)
    with quiet():
        vae = R.VQVAE(vocab_size=4096, z_channels=32, ch=32, test_mode=True, share_quant_resi=4, v_patch_nums=P256)
    vae.load_state_dict(vsd, strict=True)
    f = hashed("golden.encode.f", 0, (2, 32, 16, 16), 1.5)
    with torch.no_grad():
        idx = vae.quantize.f_to_idxBl_or_fhat(f, to_fhat=False)
        fh = vae.quantize.f_to_idxBl_or_fhat(f, to_fhat=True)
    o["f"] = f.numpy()
    for si, t in enumerate(idx):
        o[f"idx_{si}"] = t.numpy().astype(np.int16)
    o["f_hat_last"] = fh[-1].numpy()
    o["f_hat_3"] = fh[3].numpy()
    # encoder + quant_conv on a 64x64 image (latent 4x4)
    img = hashed("golden.encode.img", 0, (2, 3, 64, 64), 1.0)
    with torch.no_grad():
        o["img"] = img.numpy()
        o["feat"] = vae.quant_conv(vae.encoder(img)).numpy()
    np.savez_compressed(os.path.join(OUT, "encode.npz"), **o)
    print("wrote encode.npz", {k: v.shape for k, v in o.items()})


def width_golden():
    """north-star WIDTHS (`python oracle/make_golden.py widths` writes only this one): teacher-forced logits of the real
    reference VAR.forward (models/var.py:217-259) for (a) VAR-d16 (C=1024, H=16, 16 blocks) and (b) a 2-block model at the
    d30 width (C=1920, H=30), 256 px pyramid, B=1, with ACTIVE residual gates (gamma_bias=0.5) and a peaked head
    (init_head=1.0) so that every block moves the logits.  Stored: the first 96 logit columns of every row, the row argmax
    and |logit|max."""
    os.makedirs(OUT, exist_ok=True)
    kw = dict(gamma_bias=0.5, init_head=1.0)
    o = {}
    for name, depth, C, H, seed, lab in (("d16", 16, 1024, 16, 1, 207), ("w30", 2, 1920, 30, 2, 388)):
        with quiet():
            import models as R
            vae = R.VQVAE(vocab_size=4096, z_channels=32, ch=32, test_mode=True, share_quant_resi=4, v_patch_nums=P256)
            m = R.VAR(vae_local=vae, depth=depth, embed_dim=C, num_heads=H, attn_l2_norm=True, patch_nums=P256,
                      flash_if_available=False, fused_if_available=False).eval()
        m.load_state_dict(var_state_dict(depth, patch_nums=P256, seed=seed, tag=name, embed_dim=C, num_heads=H, **kw), strict=True)
        m.cond_drop_rate = 0.0                      # the reference drops labels even in eval (var.py:226)
        x_in = hashed(f"golden.width.{name}.x", 0, (1, 679, 32), 1.0)
        with torch.no_grad():
            logits = m(torch.tensor([lab]), x_in)
        o[f"{name}_logits_slice"] = logits[:, :, :96].numpy()
        o[f"{name}_argmax"] = logits.argmax(-1).numpy().astype(np.int16)
        o[f"{name}_absmax"] = np.float32(logits.abs().max())
        o[f"{name}_label"] = np.int64(lab)
        print(name, "absmax", float(logits.abs().max()))
    np.savez_compressed(os.path.join(OUT, "real_width.npz"), **o)
    print("wrote real_width.npz", os.path.getsize(os.path.join(OUT, "real_width.npz")) // 1024, "KiB")


def main():
    if sys.argv[1:] == ["encode"]:
        return encode_golden()
    if sys.argv[1:] == ["widths"]:
        return width_golden()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    with quiet():
        from models.helpers import sample_with_top_k_top_p_

    # ---- sampler -------------------------------------------------------------------------------------
    cases = []
    for ci, (scale, tk, tp) in enumerate([(0.05, 0, 0.0), (0.05, 900, 0.96), (3.0, 900, 0.96), (3.0, 0, 0.9), (1.0, 50, 0.5), (3.0, 600, 0.0)]):
        B, L, V = 2, 8, 4096
        si, K, cfg = 6, 10, 1.5
        lg = hashed(f"golden.sampler.{ci}", 0, (2 * B, L, V), scale)
        t = cfg * (si / (K - 1))
        mixed = (1 + t) * lg[:B] - t * lg[B:]
        g = torch.Generator().manual_seed(100 + ci)
        state = g.get_state()
        idx = sample_with_top_k_top_p_(mixed, rng=g, top_k=tk, top_p=tp, num_samples=1)[:, :, 0]
        g.set_state(state)
        noise = torch.empty(B * L, V).exponential_(generator=g)     # what multinomial drew (pin P4)
        cases.append(dict(scale=scale, top_k=tk, top_p=tp, si=si, K=K, cfg=cfg, idx=idx.numpy().astype(np.int16),
                          removed=np.packbits(torch.isinf(mixed).numpy()), noise=noise.numpy()))
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), n=len(cases),
                        **{f"{k}_{i}": np.asarray(v) for i, c in enumerate(cases) for k, v in c.items()})

    # ---- VQ next-input -------------------------------------------------------------------------------
    vsd = vqvae_state_dict(ch=32, patch_nums=P256)
    vae, _ = build_ref(P256, 1)
    load_vae(vae, vsd)
    q = vae.quantize
    f_hat = torch.zeros(1, 32, 16, 16)
    out = {}
    g = torch.Generator().manual_seed(7)
    for si, pn in enumerate(P256):
        idx = torch.randint(0, 4096, (1, pn * pn), generator=g)
        h = q.embedding(idx).transpose(1, 2).reshape(1, 32, pn, pn)
        f_hat, nm = q.get_next_autoregressive_input(si, len(P256), f_hat, h)
        out[f"idx_{si}"] = idx.numpy().astype(np.int16)
        out[f"f_hat_{si}"] = f_hat.clone().numpy()
        out[f"next_{si}"] = nm.clone().numpy()
    np.savez_compressed(os.path.join(OUT, "vq.npz"), **out)

    # ---- tiny VAR: baseline loop, teacher forcing, sd_test3, block stack --------------------------------
    vsd4 = vqvae_state_dict(ch=32, patch_nums=P4)
    kw = dict(gamma_bias=0.5, init_head=1.0)
    vae, draft = build_ref(P4, 2)
    load_vae(vae, vsd4)
    draft.load_state_dict(var_state_dict(2, patch_nums=P4, seed=1, tag="draft", **kw), strict=True)
    with quiet():
        import models as R
        target = R.VAR(vae_local=vae, depth=3, embed_dim=192, num_heads=3, attn_l2_norm=True, patch_nums=P4,
                       flash_if_available=False, fused_if_available=False).eval()
    target.load_state_dict(var_state_dict(3, patch_nums=P4, seed=2, tag="target", **kw), strict=True)
    B, lab = 2, torch.tensor([3, 977])
    o = {}
    for name, model in (("draft", draft), ("target", target)):
        for tk, tp in ((0, 0.0), (900, 0.96)):
            img = model.autoregressive_infer_cfg(B, lab, g_seed=5, cfg=1.5, top_k=tk, top_p=tp)
            o[f"{name}_img_{tk}"] = img.numpy().astype(np.float16)
    # tokens + f_hat via a hook on the quantizer
    rec = []
    orig = vae.quantize.get_next_autoregressive_input
    vae.quantize.get_next_autoregressive_input = lambda si, SN, f, h: (rec.append((si, h.clone())), orig(si, SN, f, h))[1]
    target.autoregressive_infer_cfg(B, lab, g_seed=5, cfg=1.5, top_k=900, top_p=0.96)
    cb = vae.quantize.embedding.weight
    toks = []
    for si, h in rec:
        hv = h.reshape(B, 32, -1).transpose(1, 2)                      # (B,l,32)
        toks.append(((hv.unsqueeze(2) - cb.view(1, 1, 4096, 32)).abs().sum(-1)).argmin(-1))
    for si, tkn in enumerate(toks):
        o[f"target_idx_{si}"] = tkn.numpy().astype(np.int16)
    rec.clear()
    with quiet():
        sd = R.SDVAR(draft, target)
    import models.var as RV
    _dev = torch.device
    RV.torch.device = lambda *a, **k: _dev("cpu")                      # the reference hard-codes cuda:0 (var.py:737, D11)
    try:
        with quiet():
            img = sd.sdvar_autoregressive_infer_cfg_sd_test3(B, lab, g_seed=5, cfg=1.5, top_k=900, top_p=0.96, entry_num=2, sd_mask=0)
    finally:
        RV.torch.device = _dev
    o["sd_test3_e2_img"] = img.numpy().astype(np.float16)
    toks = []
    for si, h in rec:
        hv = h.reshape(B, 32, -1).transpose(1, 2)
        toks.append(((hv.unsqueeze(2) - cb.view(1, 1, 4096, 32)).abs().sum(-1)).argmin(-1))
    for si, tkn in enumerate(toks):
        o[f"sd_test3_e2_idx_{si}"] = tkn.numpy().astype(np.int16)
    vae.quantize.get_next_autoregressive_input = orig
    # teacher-forced logits (cond_drop disabled: the reference drops labels even in eval, var.py:226)
    target.cond_drop_rate = 0.0
    x_in = hashed("golden.tf.x", 0, (B, 29, 32), 1.0)
    with torch.no_grad():
        logits = target(lab, x_in)
    o["tf_x"] = x_in.numpy()
    o["tf_logits_slice"] = logits[:, :, :64].numpy()
    o["tf_logits_argmax"] = logits.argmax(-1).numpy().astype(np.int16)
    o["tf_logits_absmax"] = np.float32(logits.abs().max())
    # block stack I/O
    xb = hashed("golden.blocks.x", 0, (2 * B, 9, 192), 1.0)
    cond = target.class_emb(torch.cat((lab, torch.full_like(lab, 1000))))
    with torch.no_grad():
        y = xb
        for blk in target.blocks:
            y = blk(x=y, cond_BD=cond, attn_bias=None)
    o["blocks_x"] = xb.numpy()
    o["blocks_y"] = y.numpy()
    np.savez_compressed(os.path.join(OUT, "tiny_var.npz"), **o)

    # ---- d16, 256 px, B=1 (configs[0] size) -------------------------------------------------------------
    vae, d16 = build_ref(P256, 16)
    load_vae(vae, vsd)
    d16.load_state_dict(var_state_dict(16, patch_nums=P256, seed=1, tag="draft"), strict=True)
    rec = []
    orig = vae.quantize.get_next_autoregressive_input
    vae.quantize.get_next_autoregressive_input = lambda si, SN, f, h: (rec.append((si, h.clone(), None)), orig(si, SN, f, h))[1]
    d16.autoregressive_infer_cfg(1, torch.tensor([207]), g_seed=0, cfg=1.5, top_k=900, top_p=0.96)
    cb = vae.quantize.embedding.weight
    o = {}
    f = torch.zeros(1, 32, 16, 16)
    for si, h, _ in rec:
        hv = h.reshape(1, 32, -1).transpose(1, 2)
        o[f"idx_{si}"] = ((hv.unsqueeze(2) - cb.view(1, 1, 4096, 32)).abs().sum(-1)).argmin(-1).numpy().astype(np.int16)
        f, _ = orig(si, 10, f, h)
    o["f_hat"] = f.numpy()
    np.savez_compressed(os.path.join(OUT, "d16.npz"), **o)
    encode_golden()
    width_golden()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)) // 1024, "KiB")


if __name__ == "__main__":
    main()
