"""Live pin of the oracle against the real reference -- only where /root/reference exists (the build container);
skipped on the GPU box, where the committed fixtures in tests/golden carry the same pins."""
import contextlib
import io
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not mounted")
P4 = (1, 2, 3, 4)


def test_oracle_baseline_loop_is_bit_identical_to_reference():
    from oracle.ref_model import RefDecoder, RefVAR, RefVQ
    from sdvar_b200.weights import var_state_dict, vqvae_state_dict
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import models as R
        vae = R.VQVAE(vocab_size=4096, z_channels=32, ch=32, test_mode=True, share_quant_resi=4, v_patch_nums=P4)
        var = R.VAR(vae_local=vae, depth=2, embed_dim=128, num_heads=2, attn_l2_norm=True, patch_nums=P4, shared_aln=True,
                    flash_if_available=False, fused_if_available=False).eval()
    vsd = vqvae_state_dict(ch=32, patch_nums=P4)
    sd = var_state_dict(2, patch_nums=P4, shared_aln=True, gamma_bias=0.5, init_head=1.0)
    r = vae.load_state_dict(vsd, strict=False)
    assert all(k.startswith(("encoder.", "quant_conv.")) for k in r.missing_keys)
    var.load_state_dict(sd, strict=True)
    lab = torch.tensor([3, 977])
    with contextlib.redirect_stdout(io.StringIO()):
        img = var.autoregressive_infer_cfg(2, lab, g_seed=5, cfg=1.5, top_k=900, top_p=0.96)
    f_hat, _ = RefVAR(sd, P4).autoregressive_infer_cfg(RefVQ(vsd, P4), 2, lab, cfg=1.5, top_k=900, top_p=0.96,
                                                       rng=torch.Generator().manual_seed(5))
    assert torch.equal(img, RefDecoder(vsd).fhat_to_img(f_hat).add_(1).mul_(0.5))     # shared-adaLN path included


def test_vqvae_state_dict_surface_and_encoder_match_reference():
    """The product VQVAE with its encode side has exactly the reference VQVAE's state-dict keys and shapes (so
    vae_ch160v4096z32.pth loads strict), and its fp32 Encoder + quant_conv reproduce the reference's features on CPU."""
    from sdvar_b200.models.vqvae import VQVAE
    from sdvar_b200.weights import hashed, vqvae_state_dict
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import models as R
        ref = R.VQVAE(vocab_size=4096, z_channels=32, ch=32, test_mode=True, share_quant_resi=4, v_patch_nums=P4)
    mine = VQVAE(vocab_size=4096, z_channels=32, ch=32, v_patch_nums=P4, with_encoder=True)
    a, b = ref.state_dict(), mine.state_dict()
    assert set(a) == set(b), sorted(set(a) ^ set(b))[:8]
    assert all(tuple(a[k].shape) == tuple(b[k].shape) for k in a)
    sd = vqvae_state_dict(ch=32, patch_nums=P4, with_encoder=True)
    ref.load_state_dict(sd, strict=True)
    mine.load_state_dict(sd, strict=True)
    img = hashed("live.encode.img", 0, (2, 3, 64, 64), 1.0)
    with torch.no_grad():
        fr = ref.quant_conv(ref.encoder(img))
        fm = mine.quant_conv(mine.encoder(img))
    assert torch.allclose(fr, fm, atol=2e-6), float((fr - fm).abs().max())
