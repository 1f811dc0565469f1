"""CPU-side checks added in round 2: the T5 text-encoder hook against vectors of the real reference wrapper, input guards
that must fire before any library call, the metric restatements against an independent implementation, packing on the host."""
import os

import numpy as np
import pytest
import torch

from oracle import textocvp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _t5_wrapper(m):
    from textocvp_b200 import modules as M, weights
    ep = M.default_exp_params(num_preds=m["num_preds"])
    ep["predictor"] = {"predictor_name": "TextOCVP_T5", "predictor_params": {
        "predictor_params": ep["predictor"]["predictor_params"]["predictor_params"],
        "fusion_params": ep["predictor"]["predictor_params"]["fusion_params"],
        "text_encoder_params": {"module": weights.t5_encoder(m["t5_seed"])}}}
    return M.setup_predictor(ep).eval()


def test_t5_hook_matches_reference_wrapper():
    """encode_text_caption with a T5 encoder (predictor_wrapper.py:100-113) is host logic + the third-party encoder, so it
    runs on the CPU: same seeded encoder, same ids / masks -> the embeddings the real reference wrapper produced."""
    from textocvp_b200 import weights
    g = torch.load(os.path.join(ROOT, "tests", "golden", "t5_hook_b2.pt"), weights_only=False)
    m = g["meta"]
    pred = _t5_wrapper(m)
    ids, mask = weights.synthetic_t5_captions(m["B"], m["L"], seed=m["cap_seed"])
    with torch.no_grad():
        text = pred.encode_text_caption(caption_tokens=ids, attn_masks=mask)
    assert text.shape == g["text_embeddings"].shape
    assert O.rel_err(text, g["text_embeddings"]) < 1e-5
    keys = sorted(k for k in pred.predictor.state_dict() if k.startswith("text_encoder."))
    assert keys == g["text_encoder_keys"]                      # strict load_state_dict contract of TextOCVP_T5
    with pytest.raises(KeyError, match="attn_masks"):
        pred.encode_text_caption(caption_tokens=ids)
    with pytest.raises(KeyError, match="caption_tokens"):
        pred.encode_text_caption(attn_masks=mask)


def test_t5_default_construction_offline():
    """Without a supplied module TextOCVP_T5 builds the t5-small architecture (random init when the pretrained weights are
    not cached), frozen, with the reference's attribute names."""
    from textocvp_b200 import modules as M
    ep = M.default_exp_params()
    ep["predictor"]["predictor_name"] = "TextOCVP_T5"
    ep["predictor"]["predictor_params"]["text_encoder_params"] = {"pretrained": False, "config": {"num_layers": 1, "vocab_size": 64}}
    pred = M.setup_predictor(ep)
    body = pred.predictor
    assert body.t5_token_dim == 512 and not any(p.requires_grad for p in body.text_encoder.parameters())
    assert "text_encoder.shared.weight" in body.state_dict()


def test_clip_guards_fire_before_any_library_call():
    from textocvp_b200 import modules as M, rollout
    savi = M.setup_model(M.default_exp_params()["model"])
    x = torch.zeros(1, 5, 3, 64, 64)
    with pytest.raises(IndexError):                             # what the reference raises for a short clip
        savi(mode="decomp", x=x, num_imgs=10, decode=False)
    with pytest.raises(ValueError):
        savi(mode="decomp", x=x, num_imgs=10, decode=False)
    with pytest.raises(ValueError, match="built for"):
        savi(mode="decomp", x=torch.zeros(1, 5, 3, 32, 32), num_imgs=5, decode=False)
    with pytest.raises(ValueError):
        savi.encode(torch.zeros(2, 3, 32, 64))
    dino = M.setup_model(M.dino_exp_params()["model"])
    with pytest.raises(IndexError):
        dino(mode="decomp", x=torch.zeros(1, 2, 81, 768), num_imgs=3, decode=False)
    with pytest.raises(ValueError, match="outside the clip"):
        rollout.frame_metrics(torch.zeros(2, 19, 3, 64, 64), torch.zeros(2, 10, 3, 64, 64), 1)
    with pytest.raises(ValueError, match="do not match"):
        rollout.frame_metrics(torch.zeros(2, 3, 3, 64, 64), torch.zeros(2, 10, 3, 32, 32), 1)


def test_tuning_is_caller_owned():
    """No process-wide knobs: the options struct lives in Python and is passed by pointer with every call."""
    from textocvp_b200 import _lib as L, ops
    assert bytes(L.TUNING) == bytes(L.Tuning())
    ops.set_gemm_mode(256)
    ops.set_tuning(decode_mode=8, no_pdl=1)
    try:
        assert (L.TUNING.gemm_mode, L.TUNING.decode_mode, L.TUNING.no_pdl) == (256, 8, 1)
        with pytest.raises(ValueError):
            ops.set_gemm_mode(7)
    finally:
        ops.set_gemm_mode(0)
        ops.set_tuning(decode_mode=0, no_pdl=0)
    assert bytes(L.TUNING) == bytes(L.Tuning())


# ---------------------------------------------------------------------------------------- metric restatements, pinned
def _ssim_cv2(x, y, ws=11, sigma=1.5, k1=0.01, k2=0.03):
    """Independent SSIM: OpenCV's Gaussian filter (the classic reference implementation of Wang et al.), float64, cropped
    to the window-valid region.  x, y: [C,H,W] numpy in [0,1]."""
    import cv2
    c1, c2 = k1 ** 2, k2 ** 2
    vals = []
    for c in range(x.shape[0]):
        a, b = x[c].astype(np.float64), y[c].astype(np.float64)
        blur = lambda t: cv2.GaussianBlur(t, (ws, ws), sigma, borderType=cv2.BORDER_REFLECT)
        h = ws // 2
        crop = lambda t: t[h:-h, h:-h]
        mx, my = blur(a), blur(b)
        sxx, syy, sxy = blur(a * a) - mx * mx, blur(b * b) - my * my, blur(a * b) - mx * my
        s = ((2 * mx * my + c1) * (2 * sxy + c2)) / ((mx * mx + my * my + c1) * (sxx + syy + c2))
        vals.append(crop(s))
    return float(np.mean(vals))


def test_ssim_restatement_against_opencv_and_known_answers():
    """The oracle's SSIM / PSNR restate piqa 1.2.2 (not installable here).  They are pinned (a) against an independent
    implementation built on OpenCV's Gaussian filter and (b) against hand-computed values."""
    g = torch.Generator().manual_seed(0)
    x = torch.rand(3, 3, 64, 64, generator=g)
    y = (x + 0.1 * torch.randn(3, 3, 64, 64, generator=g)).clamp(0, 1)
    s = O.ssim(x, y)
    for i in range(3):
        assert abs(float(s[i]) - _ssim_cv2(x[i].numpy(), y[i].numpy())) < 2e-6
    # known answers: identical images -> 1; constant images a, b -> (2ab + c1) / (a^2 + b^2 + c1) (variance terms vanish)
    assert abs(float(O.ssim(x[:1], x[:1])) - 1.0) < 1e-7
    a, b = 0.25, 0.75
    ca, cb = torch.full((1, 3, 32, 32), a), torch.full((1, 3, 32, 32), b)
    expect = (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)
    assert abs(float(O.ssim(ca, cb)) - expect) < 1e-6
    # PSNR: mse 0.01 -> 10 log10(1 / (0.01 + 1e-8)) = 19.99999566 dB
    p = O.psnr(torch.zeros(1, 3, 8, 8), torch.full((1, 3, 8, 8), 0.1))
    assert abs(float(p) - 19.99999566) < 1e-4


def test_packing_runs_on_the_host():
    """_pack arithmetic (folding, tap re-ordering, conv1(posemb)) must see host tensors: no device math library is involved."""
    from textocvp_b200 import modules as M
    savi = M.setup_model(M.default_exp_params()["model"])
    seen = []
    M._PACK_DEV[0] = torch.device("cpu")
    try:
        with torch.no_grad(), M._params_on_host(savi):
            for p in savi.parameters():
                seen.append(p.device.type)
            k, ew = {}, M.EncW()
            M._pack_encoder_convs([m.block[0] for m in savi.encoder.encoder], k, ew)
    finally:
        M._PACK_DEV[0] = None
    assert set(seen) == {"cpu"} and k["w_conv1_vp"].dtype == torch.float16
    with pytest.raises(M.L.TocvpError, match="f16 range"):
        M._f16(torch.tensor([1.0e5]))


def test_modules_copy_and_pickle_without_packed_state():
    """deepcopy / pickle drop the derived packed state (ctypes structs with device pointers cannot be pickled)."""
    import copy, io, ctypes
    from textocvp_b200 import modules as M
    ep = M.default_exp_params()
    savi, pred = M.setup_model(ep["model"]), M.setup_predictor(ep)
    # simulate a packed module: attach what a forward would have left behind
    savi._w = M.SaW(); savi._keep = {"x": torch.zeros(1)}; savi._pack_sig = ("stale",)
    savi.slot_attention._w = M.SaW(); savi.slot_attention._pack_sig = ("stale",)
    c = copy.deepcopy(savi)
    assert not hasattr(c, "_w") and not hasattr(c.slot_attention, "_pack_sig")
    assert c.slot_attention._transition is c.transition_module          # the unregistered link follows the copy
    assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), savi.state_dict().values()))
    buf = io.BytesIO()
    torch.save(savi, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    assert not hasattr(r, "_w") and r.slot_attention._transition is r.transition_module
    buf = io.BytesIO()
    torch.save(pred, buf)
    buf.seek(0)
    assert torch.load(buf, weights_only=False).predictor.token_dim == 512
