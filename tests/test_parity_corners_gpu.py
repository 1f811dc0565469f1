"""Parity corners the round-1 review asked for (all through the C ABI, on a B200):
direct transition parity, per-frame stage parity along the recurrent chain on the reference's own states, stock-init range
safety of the f16 operand path, the stand-alone sub-module forwards, the evaluator's DataParallel-wrapped drive, the
reference JSON's 576-patch geometry through decomp + rollout, and the T5 text-encoder hook."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import parity_log as PL
from oracle import textocvp_oracle as O

pytestmark = pytest.mark.gpu

STAGE_TOL = 1e-3          # north star: per-stage relative L2 error against the fp32 reference on identical stage inputs
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def models(golden_weights):
    from textocvp_b200 import modules as M
    ep = M.default_exp_params()
    savi, pred = M.setup_model(ep["model"]), M.setup_predictor(ep)
    savi.load_state_dict(golden_weights["savi_sd"], strict=True)
    body = dict(pred.predictor.state_dict())
    body.update(golden_weights["pred_sd"])
    pred.predictor.load_state_dict(body, strict=True)
    return savi.cuda().eval(), pred.cuda().eval()


@pytest.fixture(scope="module")
def dmodels(golden_dino, golden_dino_weights):
    from textocvp_b200 import modules as M
    m = golden_dino["meta"]
    ep = M.dino_exp_params(num_preds=m["num_preds"], img_size=m["img_size"], num_patches=m["N"])
    dino, pred = M.setup_model(ep["model"]), M.setup_predictor(ep)
    dino.load_state_dict(golden_dino_weights["dino_sd"], strict=True)
    body = dict(pred.predictor.state_dict())
    body.update(golden_dino_weights["pred_sd"])
    pred.predictor.load_state_dict(body, strict=True)
    return dino.cuda().eval(), pred.cuda().eval()


# ------------------------------------------------------------------------------------------------ transition (a8)
def test_transition_direct(models, golden):
    """TransformerBlock.forward (attention.py:371-396, called as savi.transition_module(slots) at SAVi.py:193) against the
    vector the real reference produced for the same input, and the fused tail of the corrector (pred_out) against it too."""
    savi, _ = models
    x = golden["sa_iter3"].cuda()
    out = savi.transition_module(x)
    PL.check(O.rel_err(out, golden["transition"]), STAGE_TOL, "transition_module(sa_iter3) vs golden transition")
    # fused path: corrector (3 iterations) + transition in one chain; its transition input is the CUDA corrector output
    sa = savi.slot_attention
    assert sa.num_iters_first == 3


def test_transition_fused_tail(models, golden, golden_weights):
    savi, _ = models
    feats = O.savi_encode(golden_weights["savi_sd"], golden_weights["videos"][:, 0], O.SAViCfg()).cuda()
    init = golden_weights["init"].cuda().contiguous()
    sa = savi.slot_attention
    out = torch.empty_like(init)
    nxt = torch.empty_like(init)
    sa.run(feats, feats.shape[1] * 128, 2, feats.shape[1], init, 3, out, 8 * 128, nxt)
    PL.check(O.rel_err(out, golden["sa_iter3"]), STAGE_TOL, "corrector(3 it) vs golden sa_iter3")
    PL.check(O.rel_err(nxt, golden["transition"]), STAGE_TOL, "fused pred_out vs golden transition")


def test_transition_direct_dino(dmodels, golden_dino):
    dino, _ = dmodels
    out = dino.transition_module(golden_dino["sa_step0"].cuda())
    PL.check(O.rel_err(out, golden_dino["transition"]), STAGE_TOL, "dino transition_module(sa_step0) vs golden")


def test_prenorm_block_raises(models):
    _, pred = models
    with pytest.raises(NotImplementedError):
        pred.predictor.predictor[0](torch.zeros(1, 8, 512, device="cuda"))


def test_chain_stage_parity_on_reference_states(models, golden, golden_weights):
    """The 20-frame decomp is a recurrence, so its end-to-end error compounds (gate 3e-3 in test_stages_gpu).  The per-stage
    contract is checked here frame by frame on IDENTICAL inputs: frame t starts from the reference's own slots of frame
    t-1 (golden slot_history), i.e. transition -> corrector(1 iteration) on the CUDA path against the reference's frame t."""
    savi, _ = models
    sd, scfg = golden_weights["savi_sd"], O.SAViCfg()
    for t in (1, 7, 19):
        feats = O.savi_encode(sd, golden_weights["videos"][:, t], scfg).cuda()
        prev = golden["slot_history"][:, t - 1].cuda().contiguous()
        cur = savi.transition_module(prev)
        out = savi.slot_attention(feats, cur, step=t)
        PL.check(O.rel_err(out, golden["slot_history"][:, t]), STAGE_TOL, f"frame {t} from the reference's frame {t - 1}")


# ------------------------------------------------------------------------------------------------ f16 operand range
def test_stock_init_rollout_is_finite(golden_weights):
    """Stock reference init (mlp_out not damped: slots grow x1.3 per step): the f16 operand path must stay finite
    over the whole 19-step rollout; PSNR against the fp32 oracle is recorded."""
    from textocvp_b200 import rollout, weights
    dev = torch.device("cuda")
    savi, pred, _ = rollout.build_models(dev, mlp_out_scale=1.0)
    videos, text, noise = weights.synthetic_inputs(2, T=20, L=32, seed=3)
    ssd, psd = weights.savi_state_dict(14), weights.predictor_state_dict(15, mlp_out_scale=1.0)
    init = ssd["initializer.slots_mu"] + ssd["initializer.slots_sigma"] * noise
    out = rollout.forward_eval(savi, pred, videos.to(dev), text.to(dev), 1, 19, init_slots=init.to(dev))
    torch.cuda.synchronize()
    for k in ("slot_history", "pred_slots", "pred_imgs", "psnr", "ssim"):
        assert torch.isfinite(out[k]).all(), k
    ref = O.rollout(ssd, psd, videos, text, init, O.SAViCfg(), O.PredCfg())
    PL.check_min(float(ref["pred_slots"][:, -1].std()), 10.0, "stock init: oracle slot std at step 19 (range reached)")
    p = O.psnr(out["pred_imgs"].cpu(), ref["pred_imgs"])
    PL.check_min(p.min(), 40.0, "stock-init rollout: frame PSNR vs fp32 oracle (dB), min")
    PL.check(O.rel_err(out["pred_slots"], ref["pred_slots"]), 5e-3, "stock-init rollout: pred_slots rel err (slot std 266)")


def test_out_of_range_activations_saturate(models):
    """Slots far outside the f16 range (1e6): every operand pack saturates at +-65504, nothing becomes inf / NaN."""
    savi, pred = models
    slots = torch.full((2, 8, 128), 1.0e6, device="cuda")
    slots[:, ::2] *= -1
    out = savi(mode="decode", slots=slots)
    for k in ("recons_imgs", "recons", "masks"):
        assert torch.isfinite(out[k]).all(), k
    sh = torch.randn(2, 3, 8, 128, device="cuda") * 3.0e5
    ps = pred(sh, text_embeddings=torch.randn(2, 32, 512, device="cuda"))
    assert torch.isfinite(ps).all()


def test_pack_rejects_weights_beyond_f16(golden_weights):
    from textocvp_b200 import _lib as L, modules as M
    ep = M.default_exp_params()
    savi = M.setup_model(ep["model"])
    sd = dict(golden_weights["savi_sd"])
    w = sd["decoder.decoder.1.block.0.weight"].clone()
    w[0, 0, 0, 0] = 1.0e5
    sd["decoder.decoder.1.block.0.weight"] = w
    savi.load_state_dict(sd, strict=True)
    savi = savi.cuda().eval()
    with pytest.raises(L.TocvpError, match="f16 range"):
        savi(mode="decode", slots=torch.zeros(1, 8, 128, device="cuda"))


# ------------------------------------------------------------------------------------------------ sub-module forwards
def _conv_ref(x, mods):
    for m in mods:
        conv = m.block[0] if hasattr(m, "block") else m
        x = F.conv2d(x, conv.weight.detach().cpu().float(), conv.bias.detach().cpu().float(), padding=conv.padding)
        if hasattr(m, "block"):
            x = torch.relu(x)
    return x


def test_conv_decoder_forward(models):
    """ConvDecoder.forward (decoders.py:111-125) on a materialised NCHW tensor against plain torch fp32 convolutions."""
    savi, _ = models
    x = torch.randn(2, 128, 64, 64, generator=torch.Generator().manual_seed(1)) * 0.5
    out = savi.decoder(x.cuda())
    assert out.shape == (2, 4, 64, 64)
    ref = _conv_ref(x, list(savi.decoder.decoder))
    PL.check(O.rel_err(out, ref), STAGE_TOL, "ConvDecoder.forward (5 convolutions) vs torch fp32")


def test_conv_decoder_forward_equals_fused_decode(models, golden):
    """decode() never builds the broadcast tensor; the materialised route through ConvDecoder.forward must agree with it."""
    savi, _ = models
    slots = golden["pred_slots"][:1, -1].cuda()
    H, W = savi.decoder_resolution
    x = slots.reshape(8, 128, 1, 1).repeat(1, 1, H, W).permute(0, 2, 3, 1).contiguous()      # SAVi.broadcast
    x = savi.decoder_pos_embedding(x).permute(0, 3, 1, 2).contiguous()
    maps = savi.decoder(x).view(1, 8, 4, H, W)
    recons, masks = maps[:, :, :3], torch.softmax(maps[:, :, 3:], dim=1)
    img = (recons * masks).sum(1)
    fused = savi(mode="decode", slots=slots)
    PL.check(O.rel_err(img, fused["recons_imgs"]), STAGE_TOL, "materialised ConvDecoder route vs fused decode")
    PL.check(O.rel_err(img, golden["dec_img"]), STAGE_TOL, "materialised ConvDecoder route vs golden dec_img")


def test_conv_encoder_forward(models, golden_weights):
    savi, _ = models
    x = golden_weights["videos"][:, 3]
    out = savi.encoder(x.cuda())
    assert out.shape == (2, 32, 64, 64)
    ref = _conv_ref(x.float(), list(savi.encoder.encoder))
    PL.check(O.rel_err(out, ref), STAGE_TOL, "SimpleConvEncoder.forward (4 convolutions) vs torch fp32")


@pytest.mark.parametrize("shape, enc_mode", [((2, 32, 128), 0), ((3, 16, 64), 0), ((1, 48, 192), 0), ((2, 64, 64), 16),
                                             ((5, 16, 96), 0)])
def test_conv_encoder_shapes(models, shape, enc_mode):
    """The encoder convolutions away from the 64 x 64 benchmark frame: the pixel-pair kernel with several 64-pixel tiles per
    row (its halo starts one pair left of a tile that is not at the image border), an odd tile count and a width that is
    not a multiple of 64 (both fall back to the 25-tap kernel), and the 25-tap kernel forced by encode_mode bit 4."""
    from textocvp_b200 import _lib as L
    savi, _ = models
    n, H, W = shape
    x = torch.rand(n, 3, H, W, generator=torch.Generator().manual_seed(11 + H + W))
    setattr(L.TUNING, "encode_mode", int(enc_mode))
    try:
        out = savi.encoder(x.cuda())
        torch.cuda.synchronize()
    finally:
        setattr(L.TUNING, "encode_mode", 0)
    assert out.shape == (n, 32, H, W)
    ref = _conv_ref(x.float(), list(savi.encoder.encoder))
    PL.check(O.rel_err(out, ref), STAGE_TOL, f"SimpleConvEncoder.forward {n}x3x{H}x{W}, encode_mode {enc_mode}, vs torch fp32")


def test_positional_modules(models):
    savi, pred = models
    x = torch.randn(3, 64, 64, 32, generator=torch.Generator().manual_seed(2))
    emb = savi.encoder_pos_embedding
    out = emb(x.cuda())
    grid = emb.grid.float()
    ref = x + F.conv2d(grid, emb.projection.weight.detach().cpu(), emb.projection.bias.detach().cpu()).permute(0, 2, 3, 1)
    PL.check(O.rel_err(out, ref), 1e-6, "SoftPositionEmbed.forward vs torch")
    pe = pred.predictor.pe
    t = torch.randn(2, 7, 8, 512, generator=torch.Generator().manual_seed(3))
    out = pe(t.cuda(), 2, 8)
    cur = torch.flip(pe.pe.detach().cpu().repeat(2, 1, 8, 1)[:, :7], dims=(1,))                # model_blocks.py:374-377
    assert torch.equal(out.cpu(), t + cur)


def test_attention_modules(models):
    """MultiHeadSelfAttention / MultiHeadCrossAttention forwards (attention.py:245-265, 303-319) against fp64 torch."""
    _, pred = models
    blk = pred.predictor.predictor[2]
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 40, 512, generator=g)

    def heads(t, h):
        B, N, E = t.shape
        return t.view(B, N, h, E // h).transpose(1, 2)

    def attn(q, k, v, h):
        q, k, v = heads(q, h), heads(k, h), heads(v, h)
        a = torch.softmax(q @ k.transpose(-1, -2) / (q.shape[-1] ** 0.5), dim=-1)
        return (a @ v).transpose(1, 2).reshape(q.shape[0], -1, h * q.shape[-1])

    sa = blk.attn
    W = lambda lin: lin.weight.detach().cpu().double()
    xd = x.double()
    ref = attn(xd @ W(sa.q).t(), xd @ W(sa.k).t(), xd @ W(sa.v).t(), 8) @ W(sa.out_projection[0]).t()
    PL.check(O.rel_err(sa(x.cuda()), ref), STAGE_TOL, "MultiHeadSelfAttention.forward vs fp64 (3 chained f16 stages)")
    ca = blk.cross_attention.cross_attn
    txt = torch.randn(3, 32, 512, generator=g)
    td = txt.double()
    ref = attn(xd @ W(ca.q).t(), td @ W(ca.k).t(), td @ W(ca.v).t(), 8) @ W(ca.out_projection).t() \
        + ca.out_projection.bias.detach().cpu().double()
    PL.check(O.rel_err(ca(txt.cuda(), x.cuda()), ref), STAGE_TOL, "MultiHeadCrossAttention.forward vs fp64")


# ------------------------------------------------------------------------------------------------ evaluator drive (b)
def test_dataparallel_wrapped_evaluator_drive(models, golden, golden_weights):
    """The reference evaluator's drive, line for line (src/base/baseEvaluator.py:142-145, 168-171 and
    src/05_evaluate_predictor.py:66-103): both models wrapped in nn.DataParallel(model.eval()), attributes read through
    .module, caption kwargs passed to BOTH calls, decode of all B*num_preds frames, clamp."""
    savi, pred = models
    decomp_model = torch.nn.DataParallel(savi.eval(), device_ids=[0]).to("cuda")
    predictor = torch.nn.DataParallel(pred.eval(), device_ids=[0]).to("cuda")
    from textocvp_b200 import weights
    tokens, lengths = weights.synthetic_captions(2, 24, seed=4)
    others = {"caption": ["a", "b"], "caption_tokens": tokens, "caption_lengths": lengths, "attn_masks": None}
    num_context, num_preds = 1, 19
    num_slots = decomp_model.module.num_slots
    slot_dim = decomp_model.module.slot_dim
    videos = golden_weights["videos"].to("cuda")
    B, L_, C, H, W = videos.shape
    torch.manual_seed(0)
    out_model = decomp_model(mode="decomp", x=videos, num_imgs=num_context + num_preds, decode=False, **others)
    slot_history = out_model["slot_history"]
    pred_slots = predictor(slot_history, **others)
    pred_slots_decode = pred_slots.reshape(B * num_preds, num_slots, slot_dim)
    out_decoder = decomp_model(mode="decode", slots=pred_slots_decode)
    pred_imgs = out_decoder.get("recons_imgs")
    pred_imgs = pred_imgs.view(B, num_preds, C, H, W).clamp(0, 1)
    targets = videos[:, num_context:num_context + num_preds].clamp(0, 1)
    assert pred_imgs.shape == targets.shape and torch.isfinite(pred_imgs).all()
    # same composition without the wrappers and with the text encoder called by hand
    text = pred.predictor.text_encoder(text=tokens.cuda(), text_length=lengths.cuda())
    torch.manual_seed(0)
    sh2 = savi(mode="decomp", x=videos, num_imgs=20, decode=False)["slot_history"]
    ps2 = pred(sh2, text_embeddings=text)
    assert torch.equal(slot_history, sh2) and torch.equal(pred_slots, ps2)


def test_multi_device_replica_is_refused(models):
    from textocvp_b200 import _lib as L
    savi, _ = models
    import copy
    rep = copy.copy(savi)
    rep._is_replica = True          # what torch.nn.parallel.replicate marks its per-device copies with
    with pytest.raises(L.TocvpError, match="one process per GPU"):
        rep._ensure_packed()


# ------------------------------------------------------------------------------------------------ 576-patch geometry
def test_dino_576_decomp_and_rollout():
    """The reference JSON's own geometry (img 336 -> 24 x 24 = 576 patch tokens, 10 slots): decomp over 3 frames + a
    3-step rollout against the oracle (slots; the image decoder of this geometry is covered by test_patch_decode_336)."""
    from textocvp_b200 import modules as M, weights
    N, T, npred = 576, 3, 3
    dsd = weights.dino_state_dict(21, img_size=336, num_patches=N, bias_scale=0.02, ln_jitter=0.05, bn_jitter=0.2)
    psd = weights.predictor_state_dict(17, mlp_out_scale=0.1, ln_jitter=0.05)
    ep = M.dino_exp_params(num_preds=npred, img_size=336, num_patches=N)
    dino, pred = M.setup_model(ep["model"]), M.setup_predictor(ep)
    dino.load_state_dict(dsd, strict=True)
    body = dict(pred.predictor.state_dict())
    body.update(psd)
    pred.predictor.load_state_dict(body, strict=True)
    dino, pred = dino.cuda().eval(), pred.cuda().eval()
    feats, text, noise = weights.synthetic_dino_inputs(2, T, N, L=16, seed=6)
    init = dsd["initializer.slots_mu"] + dsd["initializer.slots_sigma"] * noise
    dcfg = O.DinoCfg(img_size=336, num_patches=N)
    sh_ref = O.dino_decomp(dsd, feats, T, dcfg, init)
    sh = dino(mode="decomp", x=feats.cuda(), num_imgs=T, decode=False, init_slots=init.cuda())["slot_history"]
    PL.check(O.rel_err(sh[:, 0], sh_ref[:, 0]), STAGE_TOL, "N=576 decomp frame 0 (3 iterations)")
    PL.check(O.rel_err(sh, sh_ref), STAGE_TOL, "N=576 decomp, 3 frames (recurrent)")
    ps_ref = O.predictor_rollout(psd, sh_ref[:, :1], text, O.PredCfg(num_preds=npred))
    ps = pred(sh[:, :1].contiguous(), text_embeddings=text.cuda(), num_preds=npred)
    PL.check(O.rel_err(ps, ps_ref), STAGE_TOL, "N=576 3-step rollout slots")


# ------------------------------------------------------------------------------------------------ T5 hook (f2)
def test_t5_hook_rollout():
    """PredictorWrapper with TextOCVP_T5 (predictor_wrapper.py:100-113): kwargs -> HF encoder -> last_hidden_state -> rollout,
    against vectors the real reference wrapper produced with the same seeded T5 encoder (oracle/make_golden_t5.py)."""
    from textocvp_b200 import modules as M, weights
    g = torch.load(os.path.join(ROOT, "tests", "golden", "t5_hook_b2.pt"), weights_only=False)
    m = g["meta"]
    ep = M.default_exp_params(num_preds=m["num_preds"])
    ep["predictor"] = {"predictor_name": "TextOCVP_T5", "predictor_params": {
        "predictor_params": ep["predictor"]["predictor_params"]["predictor_params"],
        "fusion_params": ep["predictor"]["predictor_params"]["fusion_params"],
        "text_encoder_params": {"module": weights.t5_encoder(m["t5_seed"])}}}
    pred = M.setup_predictor(ep)
    body = dict(pred.predictor.state_dict())
    body.update(weights.predictor_state_dict(m["pred_seed"], mlp_out_scale=m["mlp_out_scale"], ln_jitter=m["ln_jitter"]))
    pred.predictor.load_state_dict(body, strict=True)
    pred = pred.cuda().eval()
    ids, mask = weights.synthetic_t5_captions(m["B"], m["L"], seed=m["cap_seed"])
    text = pred.encode_text_caption(caption_tokens=ids, attn_masks=mask)
    PL.check(O.rel_err(text, g["text_embeddings"]), 1e-4, "T5 hook: text embeddings vs reference wrapper")
    sh = torch.randn(m["B"], 1 + m["num_preds"], 8, 128, generator=torch.Generator().manual_seed(m["hist_seed"]))
    out = pred(sh.cuda(), caption_tokens=ids, attn_masks=mask, caption=["a"] * m["B"])
    PL.check(O.rel_err(out, g["pred_slots"]), STAGE_TOL, "T5 hook: 3-step rollout vs reference wrapper")
    with pytest.raises(KeyError):
        pred(sh.cuda(), caption_tokens=ids)


def test_deepcopy_and_save_after_forward(models, golden):
    """A module that has run (and therefore holds packed device state) can still be deep-copied and torch.save'd; the copy
    re-packs on its first forward and gives the same result."""
    import copy, io
    savi, _ = models
    slots = golden["pred_slots"][:1, -1].cuda()
    ref = savi(mode="decode", slots=slots)["recons_imgs"]
    c = copy.deepcopy(savi)
    assert torch.equal(c(mode="decode", slots=slots)["recons_imgs"], ref)
    buf = io.BytesIO()
    torch.save(savi, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert torch.equal(r(mode="decode", slots=slots)["recons_imgs"], ref)
