"""Evaluator composition of the rollout path (reference src/05_evaluate_predictor.py:53-104) and the metric
accumulators that are all-reduced across ranks (src/lib/metrics.py:207-212 semantics: global mean + per-frame mean).

The path shards by sequence (batch): one process per GPU, no collective inside the step; the only collective is the
final all-reduce (SUM) of the accumulator vector."""
from __future__ import annotations

from typing import Dict, Optional

import torch

import ctypes

from . import _lib as L
from . import modules as M
from . import weights


@torch.no_grad()
def frame_metrics(pred_imgs: torch.Tensor, videos: torch.Tensor, frame0: int, clamp: bool = True, want_ssim: bool = True,
                  lpips_fn=None):
    """Evaluator metrics on the device (05_evaluate_predictor.py:96-103, lib/metrics.py:181-270): pred_imgs [B,F,C,H,W]
    (unclamped) against videos[:, frame0:frame0+F] read in place; both are clamped to [0,1] inside the kernels.
    Returns dict of [B,F] tensors: mse, psnr, ssim (+ lpips when ``lpips_fn`` is given).

    LPIPS (lib/metrics.py:266-298 -> piqa.LPIPS, AlexNet) needs pretrained weights that cannot be obtained offline, so it is
    a HOOK: ``lpips_fn(pred [n,C,H,W] in [0,1], target [n,C,H,W] in [0,1]) -> [n]`` (e.g. ``piqa.LPIPS(reduction="none")``
    on the same device) is called on the clamped frames and its values flow into MetricSums like the reference's."""
    B, F_, C, H, W = pred_imgs.shape
    if videos.dim() != 5 or videos.shape[0] != B or tuple(videos.shape[2:]) != (C, H, W):
        raise ValueError(f"frame_metrics: videos {tuple(videos.shape)} do not match predictions {tuple(pred_imgs.shape)}")
    if frame0 < 0 or frame0 + F_ > videos.shape[1]:
        raise ValueError(f"frame_metrics: frames [{frame0}, {frame0 + F_}) are outside the clip of {videos.shape[1]} frames")
    pred = pred_imgs.float().contiguous()
    vid = videos.float()
    if not vid.is_contiguous():
        vid = vid.contiguous()
    n = B * F_
    mse = torch.empty(B, F_, device=pred.device, dtype=torch.float32)
    psnr = torch.empty_like(mse)
    ssim = torch.empty_like(mse) if want_ssim else None
    L.call("tocvp_frame_metrics", L.ptr(pred), L.ptr(vid), L.c_size_t(vid.stride(0)), L.c_int(F_), L.c_int(frame0),
           L.c_int(n), L.c_int(C), L.c_int(H), L.c_int(W), L.c_int(int(clamp)), L.ptr(mse), L.ptr(psnr), L.ptr(ssim),
           L.stream())
    out = {"mse": mse, "psnr": psnr, "ssim": ssim}
    if lpips_fn is not None:
        tgt = vid[:, frame0:frame0 + F_].reshape(n, C, H, W)
        p01, t01 = (pred.reshape(n, C, H, W).clamp(0, 1), tgt.clamp(0, 1)) if clamp else (pred.reshape(n, C, H, W), tgt)
        out["lpips"] = lpips_fn(p01, t01).reshape(B, F_).float()
    return out


def build_models(device, savi_seed=14, pred_seed=15, mlp_out_scale=0.1, num_context=1, num_preds=19,
                 input_buffer_size=10):
    """Named architectures (SAVi.json + TextOCVP_CustomTF.json), random init from seeds, on `device`."""
    ep = M.default_exp_params(num_context, num_preds, input_buffer_size)
    savi = M.setup_model(ep["model"])
    pred = M.setup_predictor(ep)
    savi.load_state_dict(weights.savi_state_dict(savi_seed), strict=True)
    body = dict(pred.predictor.state_dict())
    body.update(weights.predictor_state_dict(pred_seed, mlp_out_scale=mlp_out_scale))
    pred.predictor.load_state_dict(body, strict=True)
    return savi.to(device).eval(), pred.to(device).eval(), ep


def build_dino_models(device, dino_seed=16, pred_seed=17, mlp_out_scale=0.1, num_context=1, num_preds=29,
                      input_buffer_size=10, img_size=128, num_patches=81):
    """CLIPort shape (BASELINE.json configs[3]): ExtendedDINOSAUR.json with the ViT backbone replaced by synthetic patch
    features + TextOCVP_CustomTF.json, random init from seeds."""
    ep = M.dino_exp_params(num_context, num_preds, input_buffer_size, img_size, num_patches)
    dino = M.setup_model(ep["model"])
    pred = M.setup_predictor(ep)
    dino.load_state_dict(weights.dino_state_dict(dino_seed, img_size=img_size, num_patches=num_patches), strict=True)
    body = dict(pred.predictor.state_dict())
    body.update(weights.predictor_state_dict(pred_seed, mlp_out_scale=mlp_out_scale))
    pred.predictor.load_state_dict(body, strict=True)
    return dino.to(device).eval(), pred.to(device).eval(), ep


@torch.no_grad()
def forward_eval_dino(dino, pred, feats, text_embeddings, num_context, num_preds, init_slots=None, num_imgs=None,
                      only_imgs=True) -> Dict[str, torch.Tensor]:
    """Evaluator composition (05_evaluate_predictor.py:82-96) for ExtendedDINOSAUR: feats [B,T,N,F] are the frozen
    backbone's patch features; returns predicted slots and clamped predicted frames [B,num_preds,3,I,I]."""
    B = feats.shape[0]
    num_imgs = num_context + num_preds if num_imgs is None else num_imgs
    sh = dino(mode="decomp", x=feats, num_imgs=num_imgs, decode=False, init_slots=init_slots)["slot_history"]
    ps = pred(sh, text_embeddings=text_embeddings)
    dec = dino.decode(ps.reshape(B * num_preds, dino.num_slots, dino.slot_dim), only_imgs=only_imgs)
    I = dino.img_size
    imgs = dec["recons_imgs"].view(B, num_preds, 3, I, I)
    L.call("tocvp_clamp01", L.ptr(imgs), L.c_size_t(imgs.numel()), L.stream())      # 05_evaluate_predictor.py:96
    return {"slot_history": sh, "pred_slots": ps, "pred_imgs": imgs, "recons_feats": dec["recons_feats"]}


@torch.no_grad()
def forward_eval(savi, pred, videos, text_embeddings, num_context, num_preds, init_slots=None, num_imgs=None,
                 conv_events=None, only_imgs=False, want_ssim=True, clamp_output=True) -> Dict[str, torch.Tensor]:
    """videos [B,L,3,H,W] (device) -> clamped predicted frames [B,num_preds,3,H,W] + PSNR/MSE per frame.
    Mirrors Evaluator.forward_eval: decomp over num_context+num_preds frames (the reference encodes them all although
    only the seed frames are consumed when teacher_force=False), predict, decode all B*num_preds frames, clamp."""
    B, L_, C, H, W = videos.shape
    if L_ < num_context + num_preds:
        raise ValueError(f"Seq. length {L_} smaller that num_context = {num_context} + num_preds = {num_preds}")
    num_imgs = num_context + num_preds if num_imgs is None else num_imgs
    out = savi(mode="decomp", x=videos, num_imgs=num_imgs, decode=False, init_slots=init_slots)
    slot_history = out["slot_history"]
    pred_slots = pred(slot_history, text_embeddings=text_embeddings)
    dec = savi.decode(pred_slots.reshape(B * num_preds, savi.num_slots, savi.slot_dim), only_imgs=only_imgs,
                      conv_events=conv_events)
    raw = dec["recons_imgs"].view(B, num_preds, C, H, W)
    # clamp + MSE / PSNR / SSIM in the metric kernels (targets = videos[:, num_context:num_context+num_preds], in place)
    m = frame_metrics(raw, videos, num_context, clamp=True, want_ssim=want_ssim)
    if clamp_output:                                                     # 05_evaluate_predictor.py:96, in place, own kernel
        L.call("tocvp_clamp01", L.ptr(raw), L.c_size_t(raw.numel()), L.stream())
    pred_imgs = raw
    return {"slot_history": slot_history, "pred_slots": pred_slots, "pred_imgs": pred_imgs, "mse": m["mse"],
            "psnr": m["psnr"], "ssim": m["ssim"]}


class MetricSums:
    """[sum_b psnr[b,f] (F), sum_b mse[b,f] (F), sum_b ssim[b,f] (F), sum_b lpips[b,f] (F), count] in fp64; all-reduced
    once at the end.  The LPIPS slot stays zero: its AlexNet weights need network access."""

    NAMES = ("psnr", "mse", "ssim", "lpips")

    def __init__(self, num_preds, device):
        self.F = num_preds
        self.acc = torch.zeros(len(self.NAMES) * num_preds + 1, dtype=torch.float64, device=device)

    def accumulate(self, psnr, mse, ssim=None, lpips=None):
        F_ = self.F
        for k, v in enumerate((psnr, mse, ssim, lpips)):
            if v is not None:
                self.acc[k * F_:(k + 1) * F_] += v.double().sum(0)
        self.acc[-1] += psnr.shape[0]

    def all_reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
        return self

    def results(self):
        n = self.acc[-1].item()
        out = {"count": int(n)}
        for k, name in enumerate(self.NAMES):
            pf = (self.acc[k * self.F:(k + 1) * self.F] / n).tolist()
            out[f"{name}_mean"] = sum(pf) / len(pf)
            out[f"{name}_per_frame"] = pf
        return out


def shard_range(rank: int, world: int, total: int):
    """Sequences [lo, hi) owned by `rank` when `total` sequences are sharded by batch over `world` ranks
    (rank r <- sequences [r*B/G, (r+1)*B/G); remainders go to the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
