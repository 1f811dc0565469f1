"""The CPU oracle (oracle/textocvp_oracle.py) against golden vectors produced by the REAL
reference modules (oracle/make_golden.py).  This is what pins the oracle."""
import torch

from oracle import textocvp_oracle as O

TOL = 2e-5   # fp32 vs fp32, different op order only


def test_stage_vectors(golden, golden_weights):
    g, w = golden, golden_weights
    scfg, pcfg = O.SAViCfg(), O.PredCfg()
    sd = w["savi_sd"]
    feats = O.savi_encode(sd, w["videos"][:, 0], scfg)
    assert O.rel_err(feats[:, ::g["meta"]["feat_stride"]], g["encode_feats_sub"]) < TOL
    _, hist = O.slot_attention(sd, feats, w["init"], 0, scfg, return_iters=True)
    for it in (1, 2, 3):
        assert O.rel_err(hist[it - 1], g[f"sa_iter{it}"]) < TOL
    assert O.rel_err(O.slot_attention(sd, feats, w["init"], 1, scfg), g["sa_step1"]) < TOL
    assert O.rel_err(O.transition(sd, g["sa_iter3"], scfg), g["transition"]) < TOL


def test_predictor_step(golden, golden_weights):
    g, w = golden, golden_weights
    pcfg = O.PredCfg()
    sh = g["slot_history"]
    assert O.rel_err(O.predictor_step(w["pred_sd"], sh[:, :10], w["text"], pcfg), g["pred_step_n10"]) < TOL
    assert O.rel_err(O.predictor_step(w["pred_sd"], sh[:, :1], w["text"], pcfg), g["pred_step_n1"]) < TOL


def test_decode(golden, golden_weights):
    g, w = golden, golden_weights
    dec = O.savi_decode(w["savi_sd"], g["pred_slots"][:1, -1], O.SAViCfg())
    assert O.rel_err(dec["recons"], g["dec_recons"]) < TOL
    assert O.rel_err(dec["masks"], g["dec_masks"]) < TOL
    assert O.rel_err(dec["recons_imgs"], g["dec_img"]) < TOL


def test_full_rollout(golden, golden_weights):
    g, w = golden, golden_weights
    out = O.rollout(w["savi_sd"], w["pred_sd"], w["videos"], w["text"], w["init"], O.SAViCfg(), O.PredCfg())
    assert O.rel_err(out["slot_history"], g["slot_history"]) < 1e-4
    assert O.rel_err(out["pred_slots"], g["pred_slots"]) < 1e-4
    p = O.psnr(out["pred_imgs"], g["pred_imgs"])
    assert p.min() > 70.0, p.min()   # eps=1e-8 caps PSNR at 80 dB


# ---------------------------------------------------------------- CLIPort / ExtendedDINOSAUR shape (configs[3])
def test_dino_stage_vectors(golden_dino, golden_dino_weights):
    g, w = golden_dino, golden_dino_weights
    cfg = O.DinoCfg(img_size=g["meta"]["img_size"], num_patches=g["meta"]["N"])
    sd = w["dino_sd"]
    proj = O.dino_project(sd, w["feats"][:, 0])
    assert O.rel_err(proj, g["proj_feats"]) < TOL
    assert O.rel_err(O.slot_attention(sd, proj, w["init"], 0, cfg), g["sa_step0"]) < TOL
    assert O.rel_err(O.slot_attention(sd, proj, w["init"], 1, cfg), g["sa_step1"]) < TOL
    assert O.rel_err(O.transition(sd, g["sa_step0"], cfg), g["transition"]) < TOL


def test_dino_decode(golden_dino, golden_dino_weights):
    g, w = golden_dino, golden_dino_weights
    cfg = O.DinoCfg(img_size=g["meta"]["img_size"], num_patches=g["meta"]["N"])
    dec = O.mlp_patch_decode(w["dino_sd"], g["pred_slots"].reshape(-1, 10, 128), cfg)
    assert O.rel_err(dec["recons_feats"], g["pred_feats"]) < TOL
    assert O.rel_err(dec["masks"], g["pred_masks"]) < TOL
    assert O.rel_err(dec["recons_imgs"], g["pred_imgs"]) < 1e-4


def test_dino_rollout(golden_dino, golden_dino_weights):
    g, w = golden_dino, golden_dino_weights
    m = g["meta"]
    cfg = O.DinoCfg(img_size=m["img_size"], num_patches=m["N"])
    pcfg = O.PredCfg(num_context=m["num_context"], num_preds=m["num_preds"])
    out = O.dino_rollout(w["dino_sd"], w["pred_sd"], w["feats"], w["text"], w["init"], cfg, pcfg, num_imgs=m["T"])
    assert O.rel_err(out["slot_history"], g["slot_history"]) < 1e-4
    assert O.rel_err(out["pred_slots"], g["pred_slots"]) < 1e-4
    p = O.psnr(out["pred_imgs"], g["pred_imgs"].view_as(out["pred_imgs"]).clamp(0, 1))
    assert p.min() > 70.0, p.min()


def test_text_encoder_oracle_vs_reference():
    import os
    from textocvp_b200 import weights
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "text_encoder_b5.pt"), weights_only=False)
    m = g["meta"]
    sd = weights.text_encoder_state_dict(m["seed"], bias_scale=m["bias_scale"], ln_jitter=m["ln_jitter"])
    tokens, lengths = weights.synthetic_captions(m["B"], m["L"], seed=m["cap_seed"])
    assert O.rel_err(O.text_encoder(sd, tokens, lengths), g["out"]) < TOL


def test_ocvp_oracle_vs_reference():
    """VanillaTransformerPredictor / OCVPSeq / OCVPPar restatement against the real modules (one step on a 7-frame window)."""
    import os
    from textocvp_b200 import weights
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ocvp_b3.pt"), weights_only=False)
    m = g["meta"]
    for kind in ("VanillaTransformer", "OCVPSeq", "OCVPPar"):
        sd = weights.ocvp_state_dict(kind, m["seed"], bias_scale=m["bias_scale"], ln_jitter=m["ln_jitter"])
        out = O.ocvp_step(sd, g["slots"], kind, max_len=m["input_buffer_size"])
        assert O.rel_err(out, g[kind + "_step"]) < TOL, kind
