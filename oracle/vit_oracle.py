"""CPU restatement of the frozen ViT front-end (TEST INFRASTRUCTURE -- only tests/ may import this).

Follows the reference's wrapper, src/models/EncodersDecoders/timm_encoders.py:58-96 (ViTEncoder.forward / normalize_images),
around timm's VisionTransformer as the reference instantiates it (timm_encoders.py:232-254: vit_base_patch14_dinov2 --
patch 14, 768-d, 12 heads, 12 pre-norm blocks, LayerScale, qkv_bias, LayerNorm eps 1e-6, num_classes = 0).  timm is a
third-party dependency that is neither vendored in /root/reference nor installed here, so the block below restates its
published definition (timm.models.vision_transformer: PatchEmbed = Conv2d(kernel = stride = patch) -> flatten;
_pos_embed = cat(cls_token, x) + pos_embed; Block: x + ls1(attn(norm1(x))), x + ls2(mlp(norm2(x))) with
Attention = softmax(q k^T / sqrt(d_head)) v on a fused qkv Linear, Mlp = fc2(GELU(fc1(x))), LayerScale = x * gamma).
PARITY UNPINNED against timm itself; tests/test_vit_cpu.py pins the block arithmetic against torchvision's independent
EncoderBlock implementation (same pre-norm MHA + GELU MLP block, LayerScale = 1)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def vit_block(sd: Dict[str, Tensor], p: str, x: Tensor, num_heads: int, eps: float = 1e-6) -> Tensor:
    B, T, E = x.shape
    dh = E // num_heads
    h = F.layer_norm(x, (E,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, T, 3, num_heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = torch.softmax(q @ k.transpose(-2, -1) * dh ** -0.5, dim=-1)
    a = (att @ v).transpose(1, 2).reshape(B, T, E)
    a = F.linear(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    x = x + a * sd.get(p + "ls1.gamma", torch.ones(E))
    h = F.layer_norm(x, (E,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    m = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                 sd[p + "mlp.fc2.bias"])
    return x + m * sd.get(p + "ls2.gamma", torch.ones(E))


def vit_encode(sd: Dict[str, Tensor], x: Tensor, patch: int, num_heads: int, mean, std, num_blocks: Optional[int] = None,
               prefix: str = "vit_backbone.") -> Tensor:
    """ViTEncoder.forward: x [B,3,H,W] -> [B, N, E].  ``std`` is what the wrapper divides by (the reference passes the MEAN)."""
    m = torch.tensor(mean).view(1, 3, 1, 1)
    s = torch.tensor(std).view(1, 3, 1, 1)
    x = (x - m) / s
    t = F.conv2d(x, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=patch)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd[prefix + "cls_token"].expand(t.shape[0], -1, -1), t], dim=1) + sd[prefix + "pos_embed"]
    depth = len({k.split(".")[len(prefix.split(".")) - 1 + 1] for k in sd if k.startswith(prefix + "blocks.")})
    for i in range(depth if num_blocks is None else num_blocks):
        t = vit_block(sd, f"{prefix}blocks.{i}.", t, num_heads)
    return t[:, 1:]
