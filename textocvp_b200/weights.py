"""
Deterministic random-init state dicts with the reference's parameter names and shapes.

There are no checkpoints offline, so benchmarks, tests and golden vectors all use weights
drawn here (torch CPU generator, seeded) with the reference's init *distributions*
(SURVEY.md Appendix A.20): xavier-uniform matrices / zero biases for SAVi, PyTorch-default
Linear init for the predictor's cross block, mlp_in and mlp_out, pe = 512^-0.5 * randn.
The key contract is SURVEY.md Appendix B (strict ``load_state_dict`` in the reference,
src/lib/setup_model.py:224), so the dicts load into the reference modules, the oracle and
the CUDA modules alike.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor


def _xavier(g, *shape) -> Tensor:
    fan_out, fan_in = shape[0], shape[1]
    rf = 1
    for s in shape[2:]:
        rf *= s
    a = math.sqrt(6.0 / ((fan_in + fan_out) * rf))
    return (torch.rand(*shape, generator=g) * 2 - 1) * a


def _default_linear(g, out_f, in_f, bias=True):
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=g) * 2 - 1) * bound if bias else None
    return w, b


def _ln(sd, name, dim, g, jitter):
    # LayerNorm defaults are (1, 0); ``jitter`` perturbs them so that parity tests exercise
    # the affine terms instead of multiplying by exactly one.
    sd[name + ".weight"] = 1.0 + jitter * torch.randn(dim, generator=g)
    sd[name + ".bias"] = jitter * torch.randn(dim, generator=g)


def savi_state_dict(seed: int = 14, num_slots: int = 8, slot_dim: int = 128, in_channels: int = 3,
                    enc_channels=(32, 32, 32, 32), dec_channels=(64, 64, 64, 64), kernel_size: int = 5,
                    mlp_hidden: int = 256, mlp_encoder_dim: int = 128, transition_mlp: int = 512,
                    bias_scale: float = 0.0, ln_jitter: float = 0.0) -> Dict[str, Tensor]:
    """SAVi (src/models/SAVi.py) parameters.  ``bias_scale`` > 0 randomises the biases that the
    reference initialises to zero, so that tests cover the bias paths."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    D, F = slot_dim, mlp_encoder_dim

    def bias(n):
        return bias_scale * torch.randn(n, generator=g)

    lim = math.sqrt(6.0 / (1 + D))
    sd["initializer.slots_mu"] = (torch.rand(1, 1, D, generator=g) * 2 - 1) * lim
    sd["initializer.slots_sigma"] = (torch.rand(1, 1, D, generator=g) * 2 - 1) * lim
    # transition: post-norm TransformerBlock
    for n in "qkv":
        sd[f"transition_module.attn.{n}.weight"] = _xavier(g, D, D)
    sd["transition_module.attn.out_projection.0.weight"] = _xavier(g, D, D)
    sd["transition_module.mlp.0.weight"] = _xavier(g, transition_mlp, D)
    sd["transition_module.mlp.0.bias"] = bias(transition_mlp)
    sd["transition_module.mlp.2.weight"] = _xavier(g, D, transition_mlp)
    sd["transition_module.mlp.2.bias"] = bias(D)
    _ln(sd, "transition_module.layernorm_query", D, g, ln_jitter)
    _ln(sd, "transition_module.layernorm_mlp", D, g, ln_jitter)
    # encoder
    cin = in_channels
    for i, c in enumerate(enc_channels):
        sd[f"encoder.encoder.{i}.block.0.weight"] = _xavier(g, c, cin, kernel_size, kernel_size)
        sd[f"encoder.encoder.{i}.block.0.bias"] = bias(c)
        cin = c
    sd["encoder_pos_embedding.projection.weight"] = _xavier(g, cin, 4, 1, 1)
    sd["encoder_pos_embedding.projection.bias"] = bias(cin)
    _ln(sd, "encoder_mlp.0", cin, g, ln_jitter)
    sd["encoder_mlp.1.weight"] = _xavier(g, F, cin)
    sd["encoder_mlp.1.bias"] = bias(F)
    sd["encoder_mlp.3.weight"] = _xavier(g, F, F)
    sd["encoder_mlp.3.bias"] = bias(F)
    # decoder (ConvDecoder iterates hidden dims in reverse; all equal in the named config)
    sd["decoder_pos_embedding.projection.weight"] = _xavier(g, D, 4, 1, 1)
    sd["decoder_pos_embedding.projection.bias"] = bias(D)
    cin = D
    rev = list(dec_channels)[::-1]
    for i, c in enumerate(rev):
        sd[f"decoder.decoder.{i}.block.0.weight"] = _xavier(g, c, cin, kernel_size, kernel_size)
        sd[f"decoder.decoder.{i}.block.0.bias"] = bias(c)
        cin = c
    n = len(rev)
    sd[f"decoder.decoder.{n}.weight"] = _xavier(g, in_channels + 1, dec_channels[0], 3, 3)
    sd[f"decoder.decoder.{n}.bias"] = bias(in_channels + 1)
    # slot attention
    for nm in ("norm_input",):
        _ln(sd, f"slot_attention.{nm}", F, g, ln_jitter)
    for nm in ("norm_slot", "norm_mlp"):
        _ln(sd, f"slot_attention.{nm}", D, g, ln_jitter)
    sd["slot_attention.to_q.weight"] = _xavier(g, D, D)
    sd["slot_attention.to_q.bias"] = bias(D)
    sd["slot_attention.to_k.weight"] = _xavier(g, D, F)
    sd["slot_attention.to_k.bias"] = bias(D)
    sd["slot_attention.to_v.weight"] = _xavier(g, D, F)
    sd["slot_attention.to_v.bias"] = bias(D)
    sd["slot_attention.gru.weight_ih"] = _xavier(g, 3 * D, D)
    q, _ = torch.linalg.qr(torch.randn(3 * D, D, generator=g))               # orthogonal weight_hh
    sd["slot_attention.gru.weight_hh"] = q.contiguous()
    sd["slot_attention.gru.bias_ih"] = bias(3 * D)
    sd["slot_attention.gru.bias_hh"] = bias(3 * D)
    sd["slot_attention.mlp.0.weight"] = _xavier(g, mlp_hidden, D)
    sd["slot_attention.mlp.0.bias"] = bias(mlp_hidden)
    sd["slot_attention.mlp.2.weight"] = _xavier(g, D, mlp_hidden)
    sd["slot_attention.mlp.2.bias"] = bias(D)
    return sd


def predictor_state_dict(seed: int = 15, slot_dim: int = 128, token_dim: int = 512, hidden_dim: int = 2048,
                         num_layers: int = 8, cross_inner: int = 512, cross_mlp: int = 2048,
                         input_buffer_size: int = 10, mlp_out_scale: float = 0.1,
                         ln_jitter: float = 0.0) -> Dict[str, Tensor]:
    """TextOCVP body (src/models/Predictors/text_cond_OCVP.py) parameters, keys relative to the
    TextOCVP module (the PredictorWrapper adds a ``predictor.`` prefix, see Appendix B).

    ``mlp_out_scale`` = 0.1 is the stable benchmark init of SURVEY.md 8(d): stock init makes the
    19-step rollout diverge (slot std x1.3 per step) which no reduced-precision path can track
    at 40 dB.  The same weights feed the oracle and the CUDA path."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    T = token_dim
    sd["mlp_in.weight"], sd["mlp_in.bias"] = _default_linear(g, T, slot_dim)
    w, b = _default_linear(g, slot_dim, T)
    sd["mlp_out.weight"], sd["mlp_out.bias"] = w * mlp_out_scale, b * mlp_out_scale
    sd["pe.pe"] = T ** -0.5 * torch.randn(1, input_buffer_size + 1, 1, T, generator=g)
    for i in range(num_layers):
        p = f"predictor.{i}"
        for n in "qkv":
            sd[f"{p}.attn.{n}.weight"] = _xavier(g, T, T)
        sd[f"{p}.attn.out_projection.0.weight"] = _xavier(g, T, T)
        sd[f"{p}.mlp.0.weight"] = _xavier(g, hidden_dim, T)
        sd[f"{p}.mlp.0.bias"] = torch.zeros(hidden_dim)
        sd[f"{p}.mlp.2.weight"] = _xavier(g, T, hidden_dim)
        sd[f"{p}.mlp.2.bias"] = torch.zeros(T)
        _ln(sd, f"{p}.layernorm_query", T, g, ln_jitter)
        _ln(sd, f"{p}.layernorm_mlp", T, g, ln_jitter)
        c = f"{p}.cross_attention"
        for nm in ("ln_mlp", "ln_cross_att_q", "ln_cross_att_kv"):
            _ln(sd, f"{c}.{nm}", T, g, ln_jitter)
        sd[f"{c}.mlp.0.weight"], sd[f"{c}.mlp.0.bias"] = _default_linear(g, cross_mlp, T)
        sd[f"{c}.mlp.2.weight"], sd[f"{c}.mlp.2.bias"] = _default_linear(g, T, cross_mlp)
        for n in "qkv":
            sd[f"{c}.cross_attn.{n}.weight"], _ = _default_linear(g, cross_inner, T, bias=False)
        sd[f"{c}.cross_attn.out_projection.weight"], sd[f"{c}.cross_attn.out_projection.bias"] = \
            _default_linear(g, T, cross_inner)
    return sd


def synthetic_inputs(B: int, T: int = 20, L: int = 32, token_dim: int = 512, num_slots: int = 8,
                     slot_dim: int = 128, res=(64, 64), seed: int = 0):
    """videos U[0,1) [B,T,3,H,W], text embeddings N(0,1) [B,L,512], slot-init noise N(0,1) [B,S,D]
    (SURVEY.md 8(d) config 1).  The noise is injected into both paths (CPU and CUDA generators differ)."""
    g = torch.Generator().manual_seed(seed)
    videos = torch.rand(B, T, 3, res[0], res[1], generator=g)
    text = torch.randn(B, L, token_dim, generator=g)
    noise = torch.randn(B, num_slots, slot_dim, generator=g)
    return videos, text, noise


def dino_cnn_plan(img_size: int, num_patches: int, hidden_dim: int = 1024, in_dim: int = 768, num_layers: int = 4,
                  patch_size: int = 14):
    """Channel / upsampling plan of MLPPatchDecoder._build_conv_patch_decoder (reference
    src/models/EncodersDecoders/decoders.py:325-365): list of (cin, cout, upsample_after) + module indices."""
    plan, idx, cur, h = [], 0, int(num_patches ** 0.5), hidden_dim
    for i in range(num_layers):
        cin = in_dim if i == 0 else h
        if i > 0 and (i + 1) * 2 < patch_size and cur < img_size:
            h = h // 2
        up = (i + 1) * 2 < patch_size and cur < img_size
        plan.append(dict(index=idx, cin=cin, cout=h, up=up))
        idx += 2 if up else 1
        if up:
            cur *= 2
    return plan, idx, h, cur          # conv blocks, index of the final conv, its input channels, final spatial size


def dino_state_dict(seed: int = 16, num_slots: int = 10, slot_dim: int = 128, feat_dim: int = 768,
                    img_size: int = 128, num_patches: int = 81, mlp_hidden: int = 512, transition_mlp: int = 512,
                    dec_hidden: int = 1024, dec_layers: int = 4, cnn_layers: int = 4, patch_size: int = 14,
                    bias_scale: float = 0.0, ln_jitter: float = 0.0, bn_jitter: float = 0.0) -> Dict[str, Tensor]:
    """ExtendedDINOSAUR (src/models/ExtendedDINOSAUR.py, configs/models/ExtendedDINOSAUR.json) parameters without the
    frozen ViT backbone (replaced by synthetic patch features).  init_xavier_ on linear_feat_proj / transition /
    slot_attention / decoder (ExtendedDINOSAUR.py:222-236); BatchNorm running statistics default to (0, 1) and are
    perturbed by ``bn_jitter`` so that tests exercise the eval-mode BN folding."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    D = slot_dim

    def bias(n):
        return bias_scale * torch.randn(n, generator=g)

    lim = math.sqrt(6.0 / (1 + D))
    sd["initializer.slots_mu"] = (torch.rand(1, 1, D, generator=g) * 2 - 1) * lim
    sd["initializer.slots_sigma"] = (torch.rand(1, 1, D, generator=g) * 2 - 1) * lim
    for n in "qkv":
        sd[f"transition_module.attn.{n}.weight"] = _xavier(g, D, D)
    sd["transition_module.attn.out_projection.0.weight"] = _xavier(g, D, D)
    sd["transition_module.mlp.0.weight"] = _xavier(g, transition_mlp, D)
    sd["transition_module.mlp.0.bias"] = bias(transition_mlp)
    sd["transition_module.mlp.2.weight"] = _xavier(g, D, transition_mlp)
    sd["transition_module.mlp.2.bias"] = bias(D)
    _ln(sd, "transition_module.layernorm_query", D, g, ln_jitter)
    _ln(sd, "transition_module.layernorm_mlp", D, g, ln_jitter)
    _ln(sd, "linear_feat_proj.0", feat_dim, g, ln_jitter)
    sd["linear_feat_proj.1.weight"] = _xavier(g, feat_dim, feat_dim)
    sd["linear_feat_proj.1.bias"] = bias(feat_dim)
    sd["linear_feat_proj.3.weight"] = _xavier(g, D, feat_dim)
    sd["linear_feat_proj.3.bias"] = bias(D)
    # MLPPatchDecoder
    sd["decoder.pos_embed"] = torch.randn(1, 1, num_patches, D, generator=g) / (D ** 0.5)
    _ln(sd, "decoder.mlp.0", D, g, ln_jitter)
    out_dim = feat_dim + 1
    for i in range(dec_layers):
        d1 = dec_hidden if i > 0 else D
        d2 = dec_hidden if i < dec_layers - 1 else out_dim
        sd[f"decoder.mlp.{1 + 2 * i}.weight"] = _xavier(g, d2, d1)
        sd[f"decoder.mlp.{1 + 2 * i}.bias"] = bias(d2)
    plan, last_idx, last_c, _ = dino_cnn_plan(img_size, num_patches, dec_hidden, feat_dim, cnn_layers, patch_size)
    for blk in plan:
        p = f"decoder.conv_patch_decoder.{blk['index']}.block"
        c = blk["cout"]
        sd[f"{p}.0.weight"] = _xavier(g, c, blk["cin"], 3, 3)
        sd[f"{p}.0.bias"] = bias(c)
        sd[f"{p}.1.weight"] = 1.0 + bn_jitter * torch.randn(c, generator=g)
        sd[f"{p}.1.bias"] = bn_jitter * torch.randn(c, generator=g)
        sd[f"{p}.1.running_mean"] = bn_jitter * torch.randn(c, generator=g)
        sd[f"{p}.1.running_var"] = 1.0 + bn_jitter * (torch.rand(c, generator=g) - 0.5)
        sd[f"{p}.1.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    sd[f"decoder.conv_patch_decoder.{last_idx}.weight"] = _xavier(g, 3, last_c, 3, 3)
    sd[f"decoder.conv_patch_decoder.{last_idx}.bias"] = bias(3)
    # slot attention
    _ln(sd, "slot_attention.norm_input", D, g, ln_jitter)
    _ln(sd, "slot_attention.norm_slot", D, g, ln_jitter)
    _ln(sd, "slot_attention.norm_mlp", D, g, ln_jitter)
    for nm in ("to_q", "to_k", "to_v"):
        sd[f"slot_attention.{nm}.weight"] = _xavier(g, D, D)
        sd[f"slot_attention.{nm}.bias"] = bias(D)
    sd["slot_attention.gru.weight_ih"] = _xavier(g, 3 * D, D)
    q, _ = torch.linalg.qr(torch.randn(3 * D, D, generator=g))
    sd["slot_attention.gru.weight_hh"] = q.contiguous()
    sd["slot_attention.gru.bias_ih"] = bias(3 * D)
    sd["slot_attention.gru.bias_hh"] = bias(3 * D)
    sd["slot_attention.mlp.0.weight"] = _xavier(g, mlp_hidden, D)
    sd["slot_attention.mlp.0.bias"] = bias(mlp_hidden)
    sd["slot_attention.mlp.2.weight"] = _xavier(g, D, mlp_hidden)
    sd["slot_attention.mlp.2.bias"] = bias(D)
    return sd


def synthetic_dino_inputs(B: int, T: int = 30, N: int = 81, feat_dim: int = 768, L: int = 16, token_dim: int = 512,
                          num_slots: int = 10, slot_dim: int = 128, seed: int = 0):
    """ViT patch features N(0,1) [B,T,N,768] (stand-in for the frozen DINOv2 backbone), text [B,L,512], slot noise."""
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, T, N, feat_dim, generator=g)
    text = torch.randn(B, L, token_dim, generator=g)
    noise = torch.randn(B, num_slots, slot_dim, generator=g)
    return feats, text, noise


def text_encoder_state_dict(seed: int = 18, input_dim: int = 128, num_layers: int = 2, output_dim: int = 512,
                            vocab_size: int = 50, context_length: int = 50, bias_scale: float = 0.0,
                            ln_jitter: float = 0.0) -> Dict[str, Tensor]:
    """TransformerTextEncoder parameters (reference src/models/EncodersDecoders/text_encoders.py:36-87): N(0, 0.02)
    matrices and embeddings, zero biases (perturbed by ``bias_scale`` for tests), unit LayerNorms (+ ``ln_jitter``)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    D = input_dim

    def nrm(*shape):
        return 0.02 * torch.randn(*shape, generator=g)

    def bias(n):
        return bias_scale * torch.randn(n, generator=g)

    sd["token_embedding.weight"] = nrm(vocab_size, D)
    sd["position_embedding.weight"] = nrm(context_length, D)
    _ln(sd, "layer_norm", D, g, ln_jitter)
    for i in range(num_layers):
        p = f"transformer.layers.{i}"
        sd[p + ".self_attn.in_proj_weight"] = nrm(3 * D, D)
        sd[p + ".self_attn.in_proj_bias"] = bias(3 * D)
        sd[p + ".self_attn.out_proj.weight"] = nrm(D, D)
        sd[p + ".self_attn.out_proj.bias"] = bias(D)
        sd[p + ".linear1.weight"] = nrm(4 * D, D)
        sd[p + ".linear1.bias"] = bias(4 * D)
        sd[p + ".linear2.weight"] = nrm(D, 4 * D)
        sd[p + ".linear2.bias"] = bias(D)
        _ln(sd, p + ".norm1", D, g, ln_jitter)
        _ln(sd, p + ".norm2", D, g, ln_jitter)
    _ln(sd, "text_out_projection.0", D, g, ln_jitter)
    sd["text_out_projection.1.weight"] = nrm(output_dim, D)
    sd["text_out_projection.1.bias"] = bias(output_dim)
    return sd


def synthetic_captions(B: int, L: int = 24, vocab_size: int = 50, seed: int = 0):
    """Token ids [B,L] (0 = padding after the caption) and caption lengths [B] in 3..L."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(3, L + 1, (B,), generator=g)
    tokens = torch.randint(1, vocab_size, (B, L), generator=g)
    tokens = tokens * (torch.arange(L)[None] < lengths[:, None])
    return tokens.long(), lengths.long()


def ocvp_state_dict(kind: str, seed: int = 19, slot_dim: int = 128, token_dim: int = 128, hidden_dim: int = 256,
                    num_layers: int = 2, mlp_out_scale: float = 1.0, bias_scale: float = 0.0,
                    ln_jitter: float = 0.0) -> Dict[str, Tensor]:
    """VanillaTransformerPredictor / OCVPSeq / OCVPPar parameters (reference src/models/Predictors/OCVP.py): PyTorch-default
    initialisations of nn.Linear / nn.TransformerEncoderLayer (xavier in_proj, zero attention biases)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    T = token_dim
    sd["mlp_in.weight"], sd["mlp_in.bias"] = _default_linear(g, T, slot_dim)
    wo, bo = _default_linear(g, slot_dim, T)
    sd["mlp_out.weight"], sd["mlp_out.bias"] = wo * mlp_out_scale, bo * mlp_out_scale

    def enc(p):
        sd[p + ".self_attn.in_proj_weight"] = _xavier(g, 3 * T, T)
        sd[p + ".self_attn.in_proj_bias"] = bias_scale * torch.randn(3 * T, generator=g)
        sd[p + ".self_attn.out_proj.weight"], _ = _default_linear(g, T, T, bias=False)
        sd[p + ".self_attn.out_proj.bias"] = bias_scale * torch.randn(T, generator=g)
        sd[p + ".linear1.weight"], sd[p + ".linear1.bias"] = _default_linear(g, hidden_dim, T)
        sd[p + ".linear2.weight"], sd[p + ".linear2.bias"] = _default_linear(g, T, hidden_dim)
        _ln(sd, p + ".norm1", T, g, ln_jitter)
        _ln(sd, p + ".norm2", T, g, ln_jitter)

    for i in range(num_layers):
        if kind == "VanillaTransformer":
            enc(f"transformer_encoders.{i}")
        elif kind == "OCVPSeq":
            enc(f"transformer_encoders.{i}.object_encoder_block")
            enc(f"transformer_encoders.{i}.time_encoder_block")
        elif kind == "OCVPPar":        # OCVPParLayer: an encoder layer (its own self_attn is unused) + two attention modules
            p = f"transformer_encoders.{i}"
            enc(p)
            for a in ("self_attn_obj", "self_attn_time"):
                sd[f"{p}.{a}.in_proj_weight"] = _xavier(g, 3 * T, T)
                sd[f"{p}.{a}.in_proj_bias"] = bias_scale * torch.randn(3 * T, generator=g)
                sd[f"{p}.{a}.out_proj.weight"], _ = _default_linear(g, T, T, bias=False)
                sd[f"{p}.{a}.out_proj.bias"] = bias_scale * torch.randn(T, generator=g)
        else:
            raise ValueError(kind)
    return sd


T5_TEST_CONFIG = dict(vocab_size=200, d_model=512, d_kv=64, d_ff=1024, num_layers=2, num_heads=8)


def t5_encoder(seed: int = 19, config: Optional[dict] = None):
    """A ``transformers.T5EncoderModel`` of the t5-small family (d_model 512; depth / width reduced by ``config`` for
    tests) with every parameter overwritten from a seeded CPU generator, so the same weights can be rebuilt anywhere
    without depending on the library's init code: matrices 0.05 * randn, LayerNorm scales 1 + 0.05 * randn.
    Used by the T5-hook golden vectors (oracle/make_golden_t5.py) and their tests."""
    from transformers import T5Config, T5EncoderModel
    from .modules import T5_SMALL_CONFIG
    cfg = T5Config(**{**T5_SMALL_CONFIG, **(config if config is not None else T5_TEST_CONFIG)})
    enc = T5EncoderModel(cfg).eval()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(enc.named_parameters()):
            if "layer_norm" in name:
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return enc


def synthetic_t5_captions(B: int, L: int, vocab: int = 200, seed: int = 5):
    """Token ids [B, L] int64 (0 = pad, 1 = eos as in T5) and attention masks [B, L] with ragged lengths."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(L // 3, L + 1, (B,), generator=g)
    lengths[0] = L
    ids = torch.randint(2, vocab, (B, L), generator=g)
    pos = torch.arange(L)[None]
    mask = (pos < lengths[:, None]).long()
    ids = ids * mask
    ids[torch.arange(B), lengths - 1] = 1
    return ids, mask


def vit_state_dict(seed: int = 23, img_size: int = 128, patch: int = 14, embed_dim: int = 768, depth: int = 12,
                   mlp_ratio: int = 4, prefix: str = "vit_backbone.", ls_range=(0.05, 1.0)) -> Dict[str, Tensor]:
    """timm VisionTransformer parameters (the reference's vit_base_patch14_dinov2 geometry by default) from a seeded CPU
    generator: trunc-normal-like 0.02 matrices, LayerNorm (1, 0) + jitter, LayerScale gammas drawn from ``ls_range`` (trained
    DINOv2 checkpoints carry O(0.1 - 1) gammas; the 1e-5 initial value would make every block a no-op in a parity test)."""
    g = torch.Generator().manual_seed(seed)
    E, H = embed_dim, embed_dim * mlp_ratio
    N = (img_size // patch) ** 2
    sd: Dict[str, Tensor] = {}
    rn = lambda *s, std=0.02: std * torch.randn(*s, generator=g)
    sd[prefix + "cls_token"] = rn(1, 1, E, std=0.5)
    sd[prefix + "pos_embed"] = rn(1, 1 + N, E, std=0.5)
    sd[prefix + "patch_embed.proj.weight"] = rn(E, 3, patch, patch, std=0.05)
    sd[prefix + "patch_embed.proj.bias"] = rn(E, std=0.1)
    for i in range(depth):
        p = f"{prefix}blocks.{i}."
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"] = 1.0 + 0.05 * torch.randn(E, generator=g)
            sd[p + n + ".bias"] = 0.05 * torch.randn(E, generator=g)
        sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = rn(3 * E, E, std=0.04), rn(3 * E, std=0.02)
        sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = rn(E, E, std=0.04), rn(E, std=0.02)
        sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = rn(H, E, std=0.04), rn(H, std=0.02)
        sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = rn(E, H, std=0.04), rn(E, std=0.02)
        lo, hi = ls_range
        sd[p + "ls1.gamma"] = lo + (hi - lo) * torch.rand(E, generator=g)
        sd[p + "ls2.gamma"] = lo + (hi - lo) * torch.rand(E, generator=g)
    sd[prefix + "norm.weight"], sd[prefix + "norm.bias"] = torch.ones(E), torch.zeros(E)
    return sd
