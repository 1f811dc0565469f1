"""
CPU oracle for the TextOCVP rollout path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional fp32 (or fp64) restatement, in plain torch CPU ops, of the reference
algorithm for the hot path named in BASELINE.json (SAVi corrector -> text-conditioned
predictor -> spatial-broadcast decoder + compositing).  Every function takes a flat
``state_dict`` that uses the reference's own parameter names, so
the very same weights can be loaded into the reference modules, into this oracle and into
the CUDA modules of ``textocvp_b200``.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference from
/root/reference (in the build container only), runs it on seeded inputs and stores its
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those vectors.  PSNR is a restatement of piqa==1.2.2 (not installed) -> unpinned.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl
reference legs may import this file.  The product path (``textocvp_b200``) never does.

All file:line citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# configuration (the few numbers that are NOT derivable from tensor shapes)
# --------------------------------------------------------------------------------------
@dataclass
class SAViCfg:
    """src/configs/models/SAVi.json"""
    num_slots: int = 8
    slot_dim: int = 128
    num_iterations_first: int = 3
    num_iterations: int = 1
    in_channels: int = 3
    resolution: tuple = (64, 64)
    transition_heads: int = 4
    sa_eps: float = 1e-8          # attention.py:36 (epsilon)
    sa_ln_eps: float = 1e-3       # attention.py:49-51
    tf_ln_eps: float = 1e-6       # attention.py:361-362


@dataclass
class PredCfg:
    """src/configs/predictors/TextOCVP_CustomTF.json + CONFIG.py:66-71"""
    num_layers: int = 8
    n_heads: int = 8
    cross_heads: int = 8
    cross_head_dim: int = 64
    residual: bool = True
    num_context: int = 1
    num_preds: int = 19
    input_buffer_size: int = 10
    ln_eps: float = 1e-6          # attention.py:361-362, 427, 435-436


def _sub(sd: SD, prefix: str) -> SD:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _ln(x: Tensor, sd: SD, name: str, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(x: Tensor, sd: SD, name: str) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


# --------------------------------------------------------------------------------------
# positional grid  (model_utils.py:12-34, model_blocks.py:186-226)
# --------------------------------------------------------------------------------------
def build_grid(resolution) -> Tensor:
    """[1, 4, H, W] fp32 grid with channels (y, x, 1-y, 1-x), y,x in linspace(-1, 1)."""
    ranges = [np.linspace(-1.0, 1.0, num=r) for r in resolution]
    g = np.stack(np.meshgrid(*ranges, sparse=False, indexing="ij"), axis=-1)
    g = g.reshape(resolution[0], resolution[1], -1)[None].astype(np.float32)
    g = np.concatenate([g, 1.0 - g], axis=-1)                    # model_utils.py:33
    return torch.from_numpy(g).permute(0, 3, 1, 2).contiguous()  # model_blocks.py:212


def soft_pos_embed_table(sd: SD, prefix: str, resolution) -> Tensor:
    """Conv1x1(4->C) of the grid: [H, W, C].  Batch independent (model_blocks.py:215-226)."""
    grid = build_grid(resolution).to(sd[prefix + ".projection.weight"].dtype)
    emb = F.conv2d(grid, sd[prefix + ".projection.weight"], sd[prefix + ".projection.bias"])
    return emb[0].permute(1, 2, 0).contiguous()


# --------------------------------------------------------------------------------------
# SAVi.encode  (SAVi.py:226-238; encoders.py:136-159; model_blocks.py:49-108)
# --------------------------------------------------------------------------------------
def savi_encode(sd: SD, x: Tensor, cfg: SAViCfg) -> Tensor:
    """x [B,3,H,W] -> features [B, H*W, D]."""
    i = 0
    while f"encoder.encoder.{i}.block.0.weight" in sd:
        w = sd[f"encoder.encoder.{i}.block.0.weight"]
        b = sd[f"encoder.encoder.{i}.block.0.bias"]
        x = F.relu(F.conv2d(x, w, b, stride=1, padding=w.shape[-1] // 2))
        i += 1
    x = x.permute(0, 2, 3, 1)                                              # SAVi.py:232
    x = x + soft_pos_embed_table(sd, "encoder_pos_embedding", cfg.resolution)[None]
    x = torch.flatten(x, 1, 2)                                             # SAVi.py:236
    x = _ln(x, sd, "encoder_mlp.0", 1e-5)                                  # SAVi.py:116 (default eps)
    x = F.relu(_lin(x, sd, "encoder_mlp.1"))
    x = _lin(x, sd, "encoder_mlp.3")
    return x


# --------------------------------------------------------------------------------------
# SlotAttention.forward  (attention.py:67-112)
# --------------------------------------------------------------------------------------
def gru_cell(x: Tensor, h: Tensor, sd: SD, prefix: str) -> Tensor:
    """torch.nn.GRUCell semantics (gate order r, z, n) -- SURVEY Appendix A.4."""
    gi = F.linear(x, sd[prefix + ".weight_ih"], sd[prefix + ".bias_ih"])
    gh = F.linear(h, sd[prefix + ".weight_hh"], sd[prefix + ".bias_hh"])
    i_r, i_z, i_n = gi.chunk(3, -1)
    h_r, h_z, h_n = gh.chunk(3, -1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def slot_attention(sd: SD, inputs: Tensor, slots: Tensor, step: int, cfg: SAViCfg,
                   prefix: str = "slot_attention", return_iters: bool = False):
    """inputs [B,N,Df], slots [B,S,D] -> slots [B,S,D] (attention.py:67-112)."""
    B, S, D = slots.shape
    dim_feats = inputs.shape[-1]
    scale = dim_feats ** -0.5                                              # attention.py:46
    x = _ln(inputs, sd, prefix + ".norm_input", cfg.sa_ln_eps)            # :86
    k = _lin(x, sd, prefix + ".to_k")                                      # :87
    v = _lin(x, sd, prefix + ".to_v")
    iters = cfg.num_iterations_first if step == 0 else cfg.num_iterations  # :90
    hist = []
    for _ in range(iters):
        slots_prev = slots
        q = _lin(_ln(slots, sd, prefix + ".norm_slot", cfg.sa_ln_eps), sd, prefix + ".to_q")
        dots = torch.einsum("bid,bjd->bij", q, k) * scale                  # :99
        attn = dots.softmax(dim=1) + cfg.sa_eps                            # :100 softmax over SLOTS
        attn = attn / attn.sum(dim=-1, keepdim=True)                       # :102 renorm over locations
        updates = torch.einsum("bij,bjd->bid", attn, v)                    # :103
        slots = gru_cell(updates.reshape(-1, D), slots_prev.reshape(-1, D), sd, prefix + ".gru")
        slots = slots.reshape(B, S, D)
        h = _ln(slots, sd, prefix + ".norm_mlp", cfg.sa_ln_eps)
        slots = slots + _lin(F.relu(_lin(h, sd, prefix + ".mlp.0")), sd, prefix + ".mlp.2")  # :110
        hist.append(slots)
    return (slots, hist) if return_iters else slots


# --------------------------------------------------------------------------------------
# multi-head attention helpers (attention.py:183-215, 245-265, 303-319)
# --------------------------------------------------------------------------------------
def _heads(x: Tensor, h: int) -> Tensor:
    B, T, C = x.shape
    return x.view(B, T, h, C // h).transpose(1, 2)          # [B,h,T,dh]  (contiguous head chunks)


def _mha(q: Tensor, k: Tensor, v: Tensor, h: int) -> Tensor:
    qh, kh, vh = _heads(q, h), _heads(k, h), _heads(v, h)
    dh = qh.shape[-1]
    att = (qh @ kh.transpose(-1, -2)) * dh ** -0.5           # attention.py:187-188
    att = att.softmax(dim=-1)                                # no mask on this path
    out = att @ vh
    return out.transpose(1, 2).reshape(q.shape[0], q.shape[1], -1)


def self_attention(x: Tensor, sd: SD, prefix: str, heads: int) -> Tensor:
    q, k, v = (F.linear(x, sd[f"{prefix}.{n}.weight"]) for n in "qkv")    # no bias, attention.py:167-169
    return F.linear(_mha(q, k, v, heads), sd[prefix + ".out_projection.0.weight"])


# --------------------------------------------------------------------------------------
# SAVi transition: post-norm TransformerBlock (attention.py:387-395; transition_models.py:26-31)
# --------------------------------------------------------------------------------------
def transition(sd: SD, slots: Tensor, cfg: SAViCfg, prefix: str = "transition_module") -> Tensor:
    x = self_attention(slots, sd, prefix + ".attn", cfg.transition_heads)
    y = _ln(x + slots, sd, prefix + ".layernorm_query", cfg.tf_ln_eps)
    z = _lin(F.relu(_lin(y, sd, prefix + ".mlp.0")), sd, prefix + ".mlp.2")
    return _ln(z + y, sd, prefix + ".layernorm_mlp", cfg.tf_ln_eps)


# --------------------------------------------------------------------------------------
# SAVi.decode + broadcast + ConvDecoder (SAVi.py:241-275; decoders.py:52-125)
# --------------------------------------------------------------------------------------
def savi_decoder_maps(sd: SD, slots: Tensor, cfg: SAViCfg) -> Tensor:
    """slots [B',S,D] -> pre-softmax decoder maps [B'*S, 4, H, W]."""
    D = slots.shape[-1]
    H, W = cfg.resolution
    x = slots.reshape(-1, 1, 1, D).expand(-1, H, W, D)
    x = x + soft_pos_embed_table(sd, "decoder_pos_embedding", cfg.resolution)[None]
    x = x.permute(0, 3, 1, 2)
    i = 0
    while f"decoder.decoder.{i}.block.0.weight" in sd:
        w = sd[f"decoder.decoder.{i}.block.0.weight"]
        b = sd[f"decoder.decoder.{i}.block.0.bias"]
        x = F.relu(F.conv2d(x, w, b, stride=1, padding=w.shape[-1] // 2))
        i += 1
    w, b = sd[f"decoder.decoder.{i}.weight"], sd[f"decoder.decoder.{i}.bias"]
    return F.conv2d(x, w, b, stride=1, padding=1)                          # decoders.py:110-116, no activation


def savi_decode(sd: SD, slots: Tensor, cfg: SAViCfg) -> Dict[str, Tensor]:
    B = slots.shape[0]
    y = savi_decoder_maps(sd, slots, cfg)
    y = y.reshape(B, -1, cfg.in_channels + 1, y.shape[2], y.shape[3])
    recons, masks = y.split([cfg.in_channels, 1], dim=2)                   # SAVi.py:251-252 (RGB then mask)
    masks = F.softmax(masks, dim=1)                                        # over slots
    return {"recons_imgs": torch.sum(recons * masks, dim=1), "recons": recons, "masks": masks}


# --------------------------------------------------------------------------------------
# SAVi.forward_decomp (SAVi.py:152-223)
# --------------------------------------------------------------------------------------
def initial_slots(sd: SD, B: int, cfg: SAViCfg, noise: Optional[Tensor] = None) -> Tensor:
    """LearnedRandom: mu + sigma * randn  (initializers.py:87-94; sigma NOT exponentiated)."""
    mu, sigma = sd["initializer.slots_mu"], sd["initializer.slots_sigma"]
    if noise is None:
        noise = torch.randn(B, cfg.num_slots, cfg.slot_dim, dtype=mu.dtype)
    return mu + sigma * noise


def savi_decomp(sd: SD, x: Tensor, num_imgs: int, cfg: SAViCfg, init_slots: Tensor) -> Tensor:
    """x [B,T,3,H,W] -> slot_history [B,num_imgs,S,D] (decode=False branch)."""
    predicted = init_slots
    hist = []
    for t in range(num_imgs):
        feats = savi_encode(sd, x[:, t], cfg)
        slots = slot_attention(sd, feats, predicted, t, cfg)
        predicted = transition(sd, slots, cfg)
        hist.append(slots)
    return torch.stack(hist, dim=1)


# --------------------------------------------------------------------------------------
# Predictor: BaseTextOCVP.forward (text_cond_OCVP.py:79-105), AdaptedEncoderBlock
# (attention.py:504-524), TransformerDecoderBlock (attention.py:445-463)
# --------------------------------------------------------------------------------------
def predictor_block(sd: SD, x: Tensor, text: Tensor, cfg: PredCfg, p: str) -> Tensor:
    y = x + self_attention(_ln(x, sd, p + ".layernorm_query", cfg.ln_eps), sd, p + ".attn", cfg.n_heads)
    c = p + ".cross_attention"
    qe = _ln(y, sd, c + ".ln_cross_att_q", cfg.ln_eps)
    kv = _ln(text, sd, c + ".ln_cross_att_kv", cfg.ln_eps)                 # every layer, raw text (:454)
    q = F.linear(qe, sd[c + ".cross_attn.q.weight"])
    k = F.linear(kv, sd[c + ".cross_attn.k.weight"])
    v = F.linear(kv, sd[c + ".cross_attn.v.weight"])
    z = _lin(_mha(q, k, v, cfg.cross_heads), sd, c + ".cross_attn.out_projection") + y   # out-proj HAS bias
    h = _ln(z, sd, c + ".ln_mlp", cfg.ln_eps)
    z = z + _lin(F.relu(_lin(h, sd, c + ".mlp.0")), sd, c + ".mlp.2")
    h = _ln(z, sd, p + ".layernorm_mlp", cfg.ln_eps)
    return y + _lin(F.relu(_lin(h, sd, p + ".mlp.0")), sd, p + ".mlp.2")  # skip is y (:521-523)


def predictor_step(sd: SD, slots: Tensor, text: Tensor, cfg: PredCfg, return_layers: bool = False):
    """slots [B,n,S,D], text [B,L,Dt] -> next slots [B,S,D]. ``sd`` keys relative to TextOCVP module."""
    B, n, S, _ = slots.shape
    tok = _lin(slots, sd, "mlp_in")
    pe = sd["pe.pe"][:, :n]                                                # [1,n,1,Dt]
    tok = tok + torch.flip(pe, dims=(1,))                                  # model_blocks.py:375-377
    tok = tok.reshape(B, n * S, -1)
    layers = []
    for i in range(cfg.num_layers):
        tok = predictor_block(sd, tok, text, cfg, f"predictor.{i}")
        layers.append(tok)
    tok = tok.reshape(B, n, S, -1)
    out = _lin(tok[:, -1], sd, "mlp_out")
    out = out + slots[:, -1] if cfg.residual else out
    return (out, layers) if return_layers else out


def predictor_rollout(sd: SD, slot_history: Tensor, text: Tensor, cfg: PredCfg,
                      num_preds: Optional[int] = None) -> Tensor:
    """PredictorWrapper.forward (predictor_wrapper.py:50-87), teacher_force=False."""
    num_preds = cfg.num_preds if num_preds is None else num_preds
    window = slot_history[:, :cfg.num_context].clone()
    preds = []
    for _ in range(num_preds):
        cur = predictor_step(sd, window, text, cfg)
        window = torch.cat([window, cur.unsqueeze(1)], dim=1)[:, -cfg.input_buffer_size:]
        preds.append(cur)
    return torch.stack(preds, dim=1)


# --------------------------------------------------------------------------------------
# ExtendedDINOSAUR (src/models/ExtendedDINOSAUR.py:97-102, 139-214) with MLPPatchDecoder
# (src/models/EncodersDecoders/decoders.py:129-365).  The frozen ViT backbone is replaced by
# synthetic patch features [B,T,N,768] (north star), i.e. ``encoder`` = identity.
# --------------------------------------------------------------------------------------
@dataclass
class DinoCfg:
    """src/configs/models/ExtendedDINOSAUR.json (img_size / num_patches per BASELINE.json config 4)"""
    num_slots: int = 10
    slot_dim: int = 128
    num_iterations_first: int = 3
    num_iterations: int = 1
    in_channels: int = 3
    img_size: int = 128
    num_patches: int = 81
    transition_heads: int = 4
    sa_eps: float = 1e-8
    sa_ln_eps: float = 1e-3
    tf_ln_eps: float = 1e-6
    bn_eps: float = 1e-5          # nn.BatchNorm2d default (model_blocks.py:93)


def dino_project(sd: SD, feats: Tensor) -> Tensor:
    """linear_feat_proj (ExtendedDINOSAUR.py:97-102): LN(768) -> Linear -> ReLU -> Linear(->slot_dim)."""
    x = _ln(feats, sd, "linear_feat_proj.0", 1e-5)
    x = F.relu(_lin(x, sd, "linear_feat_proj.1"))
    return _lin(x, sd, "linear_feat_proj.3")


def dino_decomp(sd: SD, feats: Tensor, num_imgs: int, cfg: DinoCfg, init_slots: Tensor) -> Tensor:
    """feats [B,T,N,768] (ViT patch features) -> slot_history [B,num_imgs,S,D] (ExtendedDINOSAUR.py:176-205)."""
    predicted = init_slots
    hist = []
    for t in range(num_imgs):
        proj = dino_project(sd, feats[:, t])
        slots = slot_attention(sd, proj, predicted, t, cfg)
        predicted = transition(sd, slots, cfg)
        hist.append(slots)
    return torch.stack(hist, dim=1)


def patch_mlp(sd: SD, slots: Tensor, prefix: str = "decoder") -> Tensor:
    """broadcast + pos_embed + MLP (decoders.py:244-249, 309-322): slots [B,S,D] -> [B,S,N,out_dim]."""
    pos = sd[prefix + ".pos_embed"]                                        # [1,1,N,D]
    x = slots.unsqueeze(2) + pos                                           # decoders.py:152-199
    i = 0
    if f"{prefix}.mlp.0.weight" in sd and sd[f"{prefix}.mlp.0.weight"].dim() == 1:   # initial_layer_norm
        x = _ln(x, sd, f"{prefix}.mlp.0", 1e-5)
        i = 1
    while f"{prefix}.mlp.{i}.weight" in sd:
        x = _lin(x, sd, f"{prefix}.mlp.{i}")
        i += 1
        if f"{prefix}.mlp.{i + 1}.weight" in sd:                           # ReLU between linears only
            x = F.relu(x)
        i += 1
    return x


def patch_cnn(sd: SD, x: Tensor, cfg: DinoCfg, prefix: str = "decoder.conv_patch_decoder") -> Tensor:
    """conv_patch_decoder (decoders.py:325-365): [ConvBlock(3x3,BN,ReLU) (+ nearest x2)]* + conv3x3 -> 3.
    The upsampling rule is re-derived from the layer index exactly as the builder does (patch_size 14)."""
    idxs = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")})
    for j, i in enumerate(idxs):
        p = f"{prefix}.{i}"
        if f"{p}.block.0.weight" in sd:
            x = F.conv2d(x, sd[f"{p}.block.0.weight"], sd[f"{p}.block.0.bias"], padding=1)
            x = F.batch_norm(x, sd[f"{p}.block.1.running_mean"], sd[f"{p}.block.1.running_var"],
                             sd[f"{p}.block.1.weight"], sd[f"{p}.block.1.bias"], False, 0.0, cfg.bn_eps)
            x = F.relu(x)
            nxt = idxs[j + 1] if j + 1 < len(idxs) else i + 1
            if nxt == i + 2:                                               # an Upsample module sits between (no params)
                x = F.interpolate(x.contiguous(), scale_factor=2, mode="nearest")
        else:
            x = F.conv2d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], padding=1)
    return x


def mlp_patch_decode(sd: SD, slots: Tensor, cfg: DinoCfg) -> Dict[str, Tensor]:
    """MLPPatchDecoder.forward (decoders.py:232-282)."""
    B, S, _ = slots.shape
    dec = patch_mlp(sd, slots)
    feats, alpha = dec[..., :-1], dec[..., -1:]
    alpha = F.softmax(alpha, dim=1)                                        # over slots
    recons_feats = torch.sum(feats * alpha, dim=1)                         # [B,N,768]
    g = int(cfg.num_patches ** 0.5)
    masks = alpha.reshape(B, S, 1, g, g)
    x = recons_feats.permute(0, 2, 1).reshape(B, feats.shape[-1], g, g)
    imgs = patch_cnn(sd, x, cfg)
    if imgs.shape[-1] != cfg.img_size:
        imgs = F.interpolate(imgs, size=(cfg.img_size, cfg.img_size), mode="bilinear", align_corners=False)
    return {"recons_imgs": imgs, "recons_feats": recons_feats, "masks": masks}


def dino_rollout(dino_sd: SD, pred_sd: SD, feats: Tensor, text: Tensor, init_slots: Tensor, dcfg: DinoCfg,
                 pcfg: PredCfg, num_imgs: Optional[int] = None) -> Dict[str, Tensor]:
    """Evaluator composition for the CLIPort / ExtendedDINOSAUR shape (05_evaluate_predictor.py:82-96)."""
    B = feats.shape[0]
    num_imgs = pcfg.num_context + pcfg.num_preds if num_imgs is None else num_imgs
    sh = dino_decomp(dino_sd, feats, num_imgs, dcfg, init_slots)
    ps = predictor_rollout(pred_sd, sh, text, pcfg)
    dec = mlp_patch_decode(dino_sd, ps.reshape(B * pcfg.num_preds, dcfg.num_slots, dcfg.slot_dim), dcfg)
    imgs = dec["recons_imgs"].view(B, pcfg.num_preds, dcfg.in_channels, dcfg.img_size, dcfg.img_size).clamp(0, 1)
    return {"slot_history": sh, "pred_slots": ps, "pred_imgs": imgs,
            "pred_feats": dec["recons_feats"].view(B, pcfg.num_preds, dcfg.num_patches, -1)}


# --------------------------------------------------------------------------------------
# TransformerTextEncoder.forward (src/models/EncodersDecoders/text_encoders.py:84-124); the encoder layers are
# torch.nn.TransformerEncoderLayer(d_model, nhead, 4*d_model, activation="gelu"), post-norm (built at :42-50), eval mode.
# --------------------------------------------------------------------------------------
def text_encoder(sd: SD, text: Tensor, text_length: Tensor, num_heads: int = 4) -> Tensor:
    """text [B,L] int64 token ids (0 = padding), text_length [B] -> [B,L,output_dim].  ``sd`` keys relative to the
    TransformerTextEncoder module."""
    B, L = text.shape
    pos = torch.arange(L)[None].expand(B, L)
    x = F.embedding(text, sd["token_embedding.weight"]) + F.embedding(pos, sd["position_embedding.weight"])
    x = _ln(x, sd, "layer_norm", 1e-8)
    x = x * (text != 0).unsqueeze(-1).to(x.dtype)                          # :107-108
    pad = text_length.unsqueeze(1) < torch.ones_like(text).cumsum(dim=1)   # :110  True = masked key
    D = x.shape[-1]
    dh = D // num_heads
    i = 0
    while f"transformer.layers.{i}.self_attn.in_proj_weight" in sd:
        p = f"transformer.layers.{i}"
        qkv = F.linear(x, sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"])
        q, k, v = qkv.split(D, dim=-1)
        qh, kh, vh = (t.view(B, L, num_heads, dh).transpose(1, 2) for t in (q, k, v))
        att = (qh @ kh.transpose(-1, -2)) * dh ** -0.5
        att = att.masked_fill(pad[:, None, None, :], float("-inf")).softmax(dim=-1)
        a = (att @ vh).transpose(1, 2).reshape(B, L, D)
        a = F.linear(a, sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"])
        x = _ln(x + a, sd, p + ".norm1", 1e-5)
        h = F.linear(F.gelu(F.linear(x, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                     sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
        x = _ln(x + h, sd, p + ".norm2", 1e-5)
        i += 1
    x = _ln(x, sd, "text_out_projection.0", 1e-5)
    return F.linear(x, sd["text_out_projection.1.weight"], sd["text_out_projection.1.bias"])


# --------------------------------------------------------------------------------------
# Sibling predictors: VanillaTransformerPredictor / OCVPSeq (src/models/Predictors/OCVP.py:24-319) over pre-norm
# torch.nn.TransformerEncoderLayer blocks (ReLU, eps 1e-5, eval mode), SlotPositionalEncoding (model_blocks.py:230-290).
# --------------------------------------------------------------------------------------
def _torch_mha(sd: SD, p: str, h: Tensor, num_heads: int) -> Tensor:
    """nn.MultiheadAttention(batch_first=True)(h, h, h) with the parameters under prefix p; h [N, T, D]."""
    N, T, D = h.shape
    dh = D // num_heads
    q, k, v = F.linear(h, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]).split(D, dim=-1)
    qh, kh, vh = (t.reshape(N, T, num_heads, dh).transpose(1, 2) for t in (q, k, v))
    a = ((qh @ kh.transpose(-1, -2)) * dh ** -0.5).softmax(dim=-1) @ vh
    return F.linear(a.transpose(1, 2).reshape(N, T, D), sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def _encoder_layer_prenorm(sd: SD, p: str, x: Tensor, num_heads: int) -> Tensor:
    """x [N, T, D]: x = x + SA(LN1(x)); x = x + FFN(LN2(x))."""
    N, T, D = x.shape
    dh = D // num_heads
    h = _ln(x, sd, p + ".norm1", 1e-5)
    q, k, v = F.linear(h, sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"]).split(D, dim=-1)
    qh, kh, vh = (t.view(N, T, num_heads, dh).transpose(1, 2) for t in (q, k, v))
    a = ((qh @ kh.transpose(-1, -2)) * dh ** -0.5).softmax(dim=-1) @ vh
    a = a.transpose(1, 2).reshape(N, T, D)
    x = x + F.linear(a, sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"])
    h = _ln(x, sd, p + ".norm2", 1e-5)
    return x + F.linear(F.relu(F.linear(h, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                        sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])


def slot_positional_encoding(max_len: int, d_model: int) -> Tensor:
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def ocvp_step(sd: SD, slots: Tensor, kind: str, num_heads: int = 4, residual: bool = True, max_len: int = 10) -> Tensor:
    """slots [B,n,S,Ds] -> [B,S,Ds].  kind = "VanillaTransformer" (OCVP.py:98-135), "OCVPSeq" (OCVP.py:211-243) or "OCVPPar"
    (OCVP.py:397-432)."""
    B, n, S, _ = slots.shape
    tok = _lin(slots, sd, "mlp_in")
    D = tok.shape[-1]
    tok = tok + slot_positional_encoding(max_len, D)[:n].view(1, n, 1, D)
    i = 0
    if kind == "VanillaTransformer":
        x = tok.reshape(B, n * S, D)
        while f"transformer_encoders.{i}.self_attn.in_proj_weight" in sd:
            x = _encoder_layer_prenorm(sd, f"transformer_encoders.{i}", x, num_heads)
            i += 1
        tok = x.reshape(B, n, S, D)
    elif kind == "OCVPPar":
        # OCVPParLayer.forward / _sa_block (OCVP.py:499-546), pre-norm: x = x + MHA_obj(LN1 x) + MHA_time(LN1 x);
        # x = x + FFN(LN2 x).  Object attention runs inside each frame, time attention along each slot's history.
        while f"transformer_encoders.{i}.self_attn_obj.in_proj_weight" in sd:
            p = f"transformer_encoders.{i}"
            h = _ln(tok, sd, p + ".norm1", 1e-5)
            ho = h.reshape(B * n, S, D)
            xo = _torch_mha(sd, p + ".self_attn_obj", ho, num_heads).reshape(B, n, S, D)
            ht = h.transpose(1, 2).reshape(B * S, n, D)
            xt = _torch_mha(sd, p + ".self_attn_time", ht, num_heads).reshape(B, S, n, D).transpose(1, 2)
            tok = tok + xo + xt
            h2 = _ln(tok, sd, p + ".norm2", 1e-5)
            tok = tok + F.linear(F.relu(F.linear(h2, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                                 sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
            i += 1
    else:
        while f"transformer_encoders.{i}.object_encoder_block.self_attn.in_proj_weight" in sd:
            p = f"transformer_encoders.{i}"
            x = _encoder_layer_prenorm(sd, p + ".object_encoder_block", tok.reshape(B * n, S, D), num_heads)
            x = x.reshape(B, n, S, D).transpose(1, 2).reshape(B * S, n, D)
            x = _encoder_layer_prenorm(sd, p + ".time_encoder_block", x, num_heads)
            tok = x.reshape(B, S, n, D).transpose(1, 2)
            i += 1
    out = _lin(tok[:, -1], sd, "mlp_out")
    return out + slots[:, -1] if residual else out


# --------------------------------------------------------------------------------------
# Evaluator.forward_eval composition (05_evaluate_predictor.py:82-96) and PSNR
# --------------------------------------------------------------------------------------
def rollout(savi_sd: SD, pred_sd: SD, videos: Tensor, text: Tensor, init_slots: Tensor,
            scfg: SAViCfg, pcfg: PredCfg, num_imgs: Optional[int] = None) -> Dict[str, Tensor]:
    B = videos.shape[0]
    num_imgs = pcfg.num_context + pcfg.num_preds if num_imgs is None else num_imgs
    sh = savi_decomp(savi_sd, videos, num_imgs, scfg, init_slots)
    ps = predictor_rollout(pred_sd, sh, text, pcfg)
    dec = savi_decode(savi_sd, ps.reshape(B * pcfg.num_preds, scfg.num_slots, scfg.slot_dim), scfg)
    H, W = scfg.resolution
    imgs = dec["recons_imgs"].view(B, pcfg.num_preds, scfg.in_channels, H, W).clamp(0, 1)
    return {"slot_history": sh, "pred_slots": ps, "pred_imgs": imgs}


def psnr(x: Tensor, y: Tensor, eps: float = 1e-8) -> Tensor:
    """Per-image PSNR, value_range 1: 10*log10(1/(mse+eps)).  Restates piqa 1.2.2 (unpinned)."""
    mse = ((x - y) ** 2).flatten(-3).mean(-1)
    return 10.0 * torch.log10(1.0 / (mse + eps))


def ssim(x: Tensor, y: Tensor, window_size: int = 11, sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03,
         value_range: float = 1.0) -> Tensor:
    """Per-image SSIM [..., C, H, W] -> [...].  Restates piqa 1.2.2 `ssim` as called by src/lib/metrics.py:223-243
    (SSIM(window_size=11, sigma=1.5, n_channels=3, reduction=None)): channel-wise valid convolution with a normalised
    separable Gaussian window, cs = (2 s_xy + c2)/(s_xx + s_yy + c2), ss = (2 mu_x mu_y + c1)/(mu_x^2 + mu_y^2 + c1) * cs,
    mean over channels and positions.  piqa is not installed -> unpinned restatement."""
    lead = x.shape[:-3]
    C, H, W = x.shape[-3:]
    x, y = x.reshape(-1, C, H, W).double(), y.reshape(-1, C, H, W).double()
    g = torch.exp(-(torch.arange(window_size, dtype=torch.float64) - (window_size - 1) / 2) ** 2 / (2 * sigma ** 2))
    g = g / g.sum()
    k = (g[:, None] * g[None, :]).expand(C, 1, window_size, window_size)
    f = lambda t: F.conv2d(t, k, groups=C)
    c1, c2 = (k1 * value_range) ** 2, (k2 * value_range) ** 2
    mx, my = f(x), f(y)
    mxx, myy, mxy = mx * mx, my * my, mx * my
    sxx, syy, sxy = f(x * x) - mxx, f(y * y) - myy, f(x * y) - mxy
    cs = (2 * sxy + c2) / (sxx + syy + c2)
    ss = (2 * mxy + c1) / (mxx + myy + c1) * cs
    return ss.flatten(1).mean(-1).reshape(lead).float()


def rel_err(a: Tensor, b: Tensor) -> float:
    """Relative L2 error ||a-b|| / ||b|| in fp64 (the per-stage parity measure)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
