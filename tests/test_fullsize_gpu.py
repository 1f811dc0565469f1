"""Parity at BASELINE.json's full size (configs[2]: batch 256, 1 seed + 19 predicted frames, 20-frame decomp) through the
size-independent properties the path offers.  The oracle needs ~1 s per sequence on the CPU, so the whole batch cannot be
recomputed there; instead:

  * the two golden sequences (pinned against the REAL reference, tests/golden/cater_b2.pt) are embedded at tile-unaligned
    positions of the batch and must come out as the reference's frames (>= 40 dB, slot error <= 2e-3 after 19 recurrent steps);
  * two more sequences are re-computed by the CPU oracle as a B = 2 job and compared the same way;
  * no op of the path mixes batch elements, so the batch-256 job must be BIT-identical to (a) the same job
    with the batch order reversed and (b) the two batch-128 jobs a 2-GPU shard would run -- the multi-GPU result is then
    the single-GPU result by construction;
  * the device metric kernels agree with the oracle's PSNR / SSIM restatement on full-size frames.
"""
import pytest
import torch

from oracle import parity_log as PL
from oracle import textocvp_oracle as O

pytestmark = pytest.mark.gpu
STAGE_TOL = 1e-3         # the north star's per-stage tolerance (here: the whole 20-frame decomp stays inside ONE stage budget)
ROLLOUT_TOL = 2e-3       # slots after 19 / 29 RECURRENT predictor steps (each <= 1e-3 on identical inputs); measured 5.5-6.4e-4.
                         # The rollout contract proper is the >= 40 dB frame PSNR checked beside it (measured 68-76 dB).
B = 256
POS = (37, 201)          # where the golden sequences sit in the batch
SPOT = (5, 255)          # sequences re-computed by the CPU oracle


@pytest.fixture(scope="module")
def full(golden, golden_weights):
    from textocvp_b200 import modules as M, rollout, weights
    m = golden["meta"]
    ep = M.default_exp_params()
    savi, pred = M.setup_model(ep["model"]), M.setup_predictor(ep)
    savi.load_state_dict(golden_weights["savi_sd"], strict=True)
    body = dict(pred.predictor.state_dict())
    body.update(golden_weights["pred_sd"])
    pred.predictor.load_state_dict(body, strict=True)
    savi, pred = savi.cuda().eval(), pred.cuda().eval()
    videos, text, noise = weights.synthetic_inputs(B, m["T"], m["L"], seed=123)
    sd = golden_weights["savi_sd"]
    init = sd["initializer.slots_mu"] + sd["initializer.slots_sigma"] * noise
    for k, p in enumerate(POS):
        videos[p], text[p], init[p] = golden_weights["videos"][k], golden_weights["text"][k], golden_weights["init"][k]

    def run(idx):
        out = rollout.forward_eval(savi, pred, videos[idx].cuda(), text[idx].cuda(), 1, 19, init_slots=init[idx].cuda())
        torch.cuda.synchronize()
        return {k: v.clone() for k, v in out.items() if v is not None}

    return dict(run=run, videos=videos, text=text, init=init, out=run(torch.arange(B)))


def test_golden_sequences_inside_full_batch(full, golden):
    out = full["out"]
    assert out["pred_imgs"].shape == (B, 19, 3, 64, 64)
    idx = list(POS)
    PL.check(O.rel_err(out["slot_history"][idx], golden["slot_history"]), STAGE_TOL, "out['slot_history'][idx], golden['slot_history']")
    PL.check(O.rel_err(out["pred_slots"][idx], golden["pred_slots"]), ROLLOUT_TOL, "out['pred_slots'][idx], golden['pred_slots']")
    p = O.psnr(out["pred_imgs"][idx].cpu(), golden["pred_imgs"])
    PL.check_min(p.min(), 40.0, "frame PSNR vs reference (dB), min")


def test_oracle_spot_check_inside_full_batch(full, golden_weights):
    idx = list(SPOT)
    ref = O.rollout(golden_weights["savi_sd"], golden_weights["pred_sd"], full["videos"][idx], full["text"][idx],
                    full["init"][idx], O.SAViCfg(), O.PredCfg(num_context=1, num_preds=19))
    out = full["out"]
    PL.check(O.rel_err(out["pred_slots"][idx], ref["pred_slots"]), ROLLOUT_TOL, "out['pred_slots'][idx], ref['pred_slots']")
    p = O.psnr(out["pred_imgs"][idx].cpu(), ref["pred_imgs"])
    PL.check_min(p.min(), 40.0, "frame PSNR vs reference (dB), min")


def test_batch_order_and_sharding_are_bit_exact(full):
    out = full["out"]
    rev = full["run"](torch.arange(B - 1, -1, -1))
    for k in ("slot_history", "pred_slots", "pred_imgs", "psnr", "ssim"):
        assert torch.equal(rev[k].flip(0), out[k]), k
    for lo in (0, B // 2):                                  # the two shards of a 2-rank job (rollout.shard_range)
        sh = full["run"](torch.arange(lo, lo + B // 2))
        for k in ("pred_slots", "pred_imgs", "mse"):
            assert torch.equal(sh[k], out[k][lo:lo + B // 2]), (k, lo)


def test_device_metrics_at_full_size(full):
    out = full["out"]
    tgt = full["videos"][:, 1:20].clamp(0, 1)
    imgs = out["pred_imgs"].cpu()
    sel = [0, 100, 255]
    ps = O.psnr(imgs[sel].reshape(-1, 3, 64, 64), tgt[sel].reshape(-1, 3, 64, 64)).view(len(sel), 19)
    ss = O.ssim(imgs[sel].reshape(-1, 3, 64, 64), tgt[sel].reshape(-1, 3, 64, 64)).view(len(sel), 19)
    assert (out["psnr"][sel].cpu() - ps).abs().max() < 1e-3
    assert (out["ssim"][sel].cpu() - ss).abs().max() < 1e-4
    assert torch.isfinite(out["pred_imgs"]).all() and torch.isfinite(out["pred_slots"]).all()


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE configs[3]: CLIPort shape with ExtendedDINOSAUR (128 x 128 frames, 81 ViT patch features, MLP patch decoder),
# batch 128, 1 seed + 29 predicted frames.
# ----------------------------------------------------------------------------------------------------------------------
DB, DPREDS = 128, 29
DSPOT = (3, 127)


@pytest.fixture(scope="module")
def full_dino(golden_dino, golden_dino_weights):
    from textocvp_b200 import modules as M, rollout, weights
    m = golden_dino["meta"]
    ep = M.dino_exp_params(num_context=1, num_preds=DPREDS, img_size=m["img_size"], num_patches=m["N"])
    dino, pred = M.setup_model(ep["model"]), M.setup_predictor(ep)
    dino.load_state_dict(golden_dino_weights["dino_sd"], strict=True)
    body = dict(pred.predictor.state_dict())
    body.update(golden_dino_weights["pred_sd"])
    pred.predictor.load_state_dict(body, strict=True)
    dino, pred = dino.cuda().eval(), pred.cuda().eval()
    feats, text, noise = weights.synthetic_dino_inputs(DB, 1 + DPREDS, m["N"], L=m["L"], seed=77)
    sd = golden_dino_weights["dino_sd"]
    init = sd["initializer.slots_mu"] + sd["initializer.slots_sigma"] * noise

    def run(idx):
        out = rollout.forward_eval_dino(dino, pred, feats[idx].cuda(), text[idx].cuda(), 1, DPREDS,
                                        init_slots=init[idx].cuda())
        torch.cuda.synchronize()
        return {k: v.clone() for k, v in out.items() if v is not None}

    return dict(run=run, feats=feats, text=text, init=init, out=run(torch.arange(DB)), meta=m)


def test_cliport_oracle_spot_check_inside_full_batch(full_dino, golden_dino_weights):
    m = full_dino["meta"]
    idx = list(DSPOT)
    ref = O.dino_rollout(golden_dino_weights["dino_sd"], golden_dino_weights["pred_sd"], full_dino["feats"][idx],
                         full_dino["text"][idx], full_dino["init"][idx],
                         O.DinoCfg(img_size=m["img_size"], num_patches=m["N"]), O.PredCfg(num_context=1, num_preds=DPREDS))
    out = full_dino["out"]
    assert out["pred_imgs"].shape == (DB, DPREDS, 3, m["img_size"], m["img_size"])
    PL.check(O.rel_err(out["slot_history"][idx], ref["slot_history"]), STAGE_TOL, "out['slot_history'][idx], ref['slot_history']")
    PL.check(O.rel_err(out["pred_slots"][idx], ref["pred_slots"]), ROLLOUT_TOL, "out['pred_slots'][idx], ref['pred_slots']")
    p = O.psnr(out["pred_imgs"][idx].cpu(), ref["pred_imgs"])
    PL.check_min(p.min(), 40.0, "frame PSNR vs reference (dB), min")


def test_cliport_batch_order_and_sharding(full_dino):
    """Batch order: bit-exact.  Sharding into two batch-64 jobs: the corrector chain is bit-exact; in the predictor the
    kernel choice depends on the row count (the LayerNorm-folded CTA-pair GEMMs need >= 1024 rows = B * n * 10, which a
    batch of 64 only reaches from the second step on), so the first step runs through the unfolded kernels and the rollouts
    agree to rounding (measured ~1e-4 per step), not bit for bit -- hence a tolerance here, unlike the CATER shape where
    both batch sizes take the same kernels."""
    out = full_dino["out"]
    rev = full_dino["run"](torch.arange(DB - 1, -1, -1))
    for k in ("slot_history", "pred_slots", "pred_imgs"):
        assert torch.equal(rev[k].flip(0), out[k]), k
    for lo in (0, DB // 2):
        sh = full_dino["run"](torch.arange(lo, lo + DB // 2))
        assert torch.equal(sh["slot_history"], out["slot_history"][lo:lo + DB // 2]), lo
        PL.check(O.rel_err(sh["pred_slots"], out["pred_slots"][lo:lo + DB // 2]), STAGE_TOL, "sh['pred_slots'], out['pred_slots'][lo:lo + DB // 2]")
        p = O.psnr(sh["pred_imgs"].cpu(), out["pred_imgs"][lo:lo + DB // 2].cpu())
        PL.check_min(p.min(), 50.0, "frame PSNR vs reference (dB), min")
