"""The ViT front-end's CPU restatement (oracle/vit_oracle.py) against an INDEPENDENT implementation of the same block:
torchvision's EncoderBlock (pre-norm nn.MultiheadAttention + GELU MLP).  timm itself is not installable here (parity against
timm: unpinned); what this pins is the block arithmetic -- fused-qkv head split, softmax scaling, residual wiring, GELU --
and the key contract of the CUDA module."""
import torch

from oracle import vit_oracle as VO


def test_block_matches_torchvision_encoder_block():
    from torchvision.models.vision_transformer import EncoderBlock
    from textocvp_b200 import weights
    E, H, heads = 128, 512, 2
    sd = weights.vit_state_dict(5, img_size=28, embed_dim=E, depth=1, ls_range=(1.0, 1.0))
    p = "vit_backbone.blocks.0."
    blk = EncoderBlock(num_heads=heads, hidden_dim=E, mlp_dim=H, dropout=0.0, attention_dropout=0.0,
                       norm_layer=lambda d: torch.nn.LayerNorm(d, eps=1e-6)).eval()
    tv = {"ln_1.weight": sd[p + "norm1.weight"], "ln_1.bias": sd[p + "norm1.bias"],
          "self_attention.in_proj_weight": sd[p + "attn.qkv.weight"], "self_attention.in_proj_bias": sd[p + "attn.qkv.bias"],
          "self_attention.out_proj.weight": sd[p + "attn.proj.weight"], "self_attention.out_proj.bias": sd[p + "attn.proj.bias"],
          "ln_2.weight": sd[p + "norm2.weight"], "ln_2.bias": sd[p + "norm2.bias"],
          "mlp.0.weight": sd[p + "mlp.fc1.weight"], "mlp.0.bias": sd[p + "mlp.fc1.bias"],
          "mlp.3.weight": sd[p + "mlp.fc2.weight"], "mlp.3.bias": sd[p + "mlp.fc2.bias"]}
    blk.load_state_dict(tv, strict=True)
    x = torch.randn(3, 5, E, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = blk(x)
    out = VO.vit_block(sd, p, x, heads)
    assert (out - ref).abs().max() < 1e-5


def test_oracle_patch_embedding_and_tokens():
    """Patch embedding = non-overlapping conv; the class token is dropped; positions are added per token."""
    from textocvp_b200 import weights
    sd = weights.vit_state_dict(6, img_size=28, embed_dim=64, depth=0)
    x = torch.rand(2, 3, 28, 28, generator=torch.Generator().manual_seed(2))
    out = VO.vit_encode(sd, x, 14, 1, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5))
    assert out.shape == (2, 4, 64)
    xn = (x - 0.5) / 0.5
    w, b = sd["vit_backbone.patch_embed.proj.weight"], sd["vit_backbone.patch_embed.proj.bias"]
    manual = (xn[0, :, 14:28, 0:14] * w[7]).sum() + b[7] + sd["vit_backbone.pos_embed"][0, 1 + 2, 7]   # patch (gy=1, gx=0)
    assert abs(float(out[0, 2, 7] - manual)) < 1e-4


def test_module_key_contract_and_packing():
    """state_dict keys of ExtendedDINOSAUR(build_backbone=True) are the reference's encoder.vit_backbone.* names."""
    from textocvp_b200 import modules as M, weights
    ep = M.dino_exp_params(img_size=28, num_patches=4)
    ep["model"]["model_params"]["build_backbone"] = True
    dino = M.setup_model(ep["model"])
    sd = weights.vit_state_dict(7, img_size=28)
    mine = {k for k in dino.state_dict() if k.startswith("encoder.")}
    assert mine == {"encoder." + k for k in sd}
    dino.encoder.load_state_dict(sd, strict=True)
    assert not any(p.requires_grad for p in dino.encoder.parameters())
    # the reference divides by the mean (timm_encoders.py:54-56); reference_std=False uses the ImageNet std
    assert torch.equal(dino.encoder.std, dino.encoder.mean)
    enc = M.get_vit_encoder({"encoder_name": "vit_base_patch14_dinov2", "encoder_params": {}}, 28, reference_std=False)
    assert abs(float(enc.std.reshape(3)[0]) - 0.229) < 1e-6
