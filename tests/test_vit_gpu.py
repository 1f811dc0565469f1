"""Frozen ViT front-end on the CUDA path against the CPU restatement oracle/vit_oracle.py.
Parity against timm itself is UNPINNED (timm is not installable here); the restatement is pinned against torchvision's
independent block implementation in tests/test_vit_cpu.py."""
import pytest
import torch

from oracle import parity_log as PL
from oracle import textocvp_oracle as O
from oracle import vit_oracle as VO

pytestmark = pytest.mark.gpu
STAGE_TOL = 1e-3
MEAN = (0.485, 0.456, 0.406)


def _encoder(img_size, sd, num_blocks=None):
    from textocvp_b200 import modules as M
    enc = M.get_vit_encoder({"encoder_name": "vit_base_patch14_dinov2", "encoder_params": {"num_blocks": num_blocks}}, img_size)
    full = dict(enc.state_dict())
    full.update({k: v for k, v in sd.items() if k in full})
    enc.load_state_dict(full, strict=True)
    return enc.cuda().eval()


def test_vit_single_block_and_patch_embedding():
    """One stage at a time on identical inputs: patch embedding + positions (0 blocks) and one transformer block."""
    from textocvp_b200 import weights
    sd = weights.vit_state_dict(23, img_size=128)
    x = torch.rand(3, 3, 128, 128, generator=torch.Generator().manual_seed(0))
    for nb in (0, 1):
        enc = _encoder(128, sd, num_blocks=nb)
        out = enc(x.cuda())
        assert out.shape == (3, 81, 768)
        ref = VO.vit_encode(sd, x, 14, 12, MEAN, MEAN, num_blocks=nb)
        PL.check(O.rel_err(out, ref), STAGE_TOL, f"ViT-B/14 @128, {nb} block(s) vs restatement")


def test_vit_full_depth_128():
    """BASELINE configs[3] geometry: 128 x 128 -> 81 patch tokens (+ class token = 82 keys: the short attention kernel),
    all 12 blocks, 5-dim input [B, T, 3, H, W]."""
    from textocvp_b200 import weights
    sd = weights.vit_state_dict(23, img_size=128)
    enc = _encoder(128, sd)
    x = torch.rand(2, 2, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    out = enc(x.cuda())
    assert out.shape == (2, 2, 81, 768)
    ref = VO.vit_encode(sd, x.reshape(4, 3, 128, 128), 14, 12, MEAN, MEAN).reshape(2, 2, 81, 768)
    PL.check(O.rel_err(out, ref), STAGE_TOL, "ViT-B/14 @128, 12 chained blocks vs restatement")


def test_vit_336_long_attention():
    """The reference JSON's geometry: 336 x 336 -> 576 patch tokens + class token = 577 keys (streaming attention kernel)."""
    from textocvp_b200 import weights
    sd = weights.vit_state_dict(24, img_size=336, depth=3)
    enc = _encoder(336, sd, num_blocks=3)
    x = torch.rand(2, 3, 336, 336, generator=torch.Generator().manual_seed(2))
    out = enc(x.cuda())
    assert out.shape == (2, 576, 768)
    ref = VO.vit_encode(sd, x, 14, 12, MEAN, MEAN, num_blocks=3)
    PL.check(O.rel_err(out, ref), STAGE_TOL, "ViT-B/14 @336 (577 tokens), 3 blocks vs restatement")


def test_long_attention_kernel_matches_short():
    """The streaming attention kernel against fp64 torch on a ragged length (Tk = 200: 3 full key blocks + 8 keys)."""
    from textocvp_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, T, H = 2, 200, 3
    qkv = (torch.randn(B * T, 3 * H * 64, generator=g) * 0.7).cuda().half()
    out = ops.mha_f16(qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:], B, T, T, H).float().cpu()
    q, k, v = (qkv.float().cpu().double().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v).transpose(1, 2).reshape(B * T, H * 64)
    PL.check(O.rel_err(out, ref), STAGE_TOL, "streaming attention (200 keys) vs fp64")


def test_dinosaur_with_backbone_images_to_slots():
    """ExtendedDINOSAUR(build_backbone=True): images -> ViT -> projection -> corrector chain, against the oracle chain."""
    from textocvp_b200 import modules as M, weights
    vsd = weights.vit_state_dict(25, img_size=128)
    dsd = weights.dino_state_dict(16, img_size=128, num_patches=81, bias_scale=0.02, ln_jitter=0.05, bn_jitter=0.2)
    ep = M.dino_exp_params(num_preds=2)
    ep["model"]["model_params"]["build_backbone"] = True
    dino = M.setup_model(ep["model"])
    full = dict(dino.state_dict())
    full.update(dsd)
    full.update({"encoder." + k: v for k, v in vsd.items()})
    dino.load_state_dict(full, strict=True)
    dino = dino.cuda().eval()
    g = torch.Generator().manual_seed(4)
    imgs = torch.rand(2, 3, 3, 128, 128, generator=g)
    noise = torch.randn(2, 10, 128, generator=g)
    init = dsd["initializer.slots_mu"] + dsd["initializer.slots_sigma"] * noise
    out = dino(mode="decomp", x=imgs.cuda(), num_imgs=3, decode=False, init_slots=init.cuda())
    feats = VO.vit_encode(vsd, imgs.reshape(6, 3, 128, 128), 14, 12, MEAN, MEAN).reshape(2, 3, 81, 768)
    PL.check(O.rel_err(out["encoded_img_feats"], feats), STAGE_TOL, "backbone features inside forward_decomp")
    ref = O.dino_decomp(dsd, feats, 3, O.DinoCfg(), init)
    PL.check(O.rel_err(out["slot_history"], ref), STAGE_TOL, "images -> slot_history (ViT + projection + corrector chain)")
