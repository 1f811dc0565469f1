"""
Generate the golden vectors under tests/golden/ by running the REAL reference modules
(imported read-only from /root/reference, see oracle/ref_import.py) on seeded inputs.

TEST INFRASTRUCTURE -- run in the build container only:

    python -m oracle.make_golden

The weights are not stored: they are regenerated from seeds by ``textocvp_b200.weights`` (torch CPU
generator, deterministic for the pinned torch build) and loaded strictly into the reference
modules.  Stored: meta (seeds, config), and the reference's outputs per stage and for the
whole 19-step rollout (sub-sampled where large).
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_import  # noqa: E402
from textocvp_b200 import weights  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

META = dict(B=2, T=20, L=32, savi_seed=14, pred_seed=15, input_seed=0,
            bias_scale=0.02, ln_jitter=0.05, mlp_out_scale=0.1, num_preds=19, num_context=1,
            feat_stride=16)


def main():
    torch.set_num_threads(os.cpu_count())
    m = META
    savi_sd = weights.savi_state_dict(m["savi_seed"], bias_scale=m["bias_scale"], ln_jitter=m["ln_jitter"])
    pred_sd = weights.predictor_state_dict(m["pred_seed"], mlp_out_scale=m["mlp_out_scale"],
                                           ln_jitter=m["ln_jitter"])
    videos, text, noise = weights.synthetic_inputs(m["B"], m["T"], m["L"], seed=m["input_seed"])
    savi, pred = ref_import.build_reference(num_preds=m["num_preds"], num_context=m["num_context"])
    ref_import.load_weights(savi, pred, savi_sd, pred_sd)
    pred.encode_text_caption = lambda **kw: text
    B = m["B"]
    out = {"meta": m}
    with torch.no_grad():
        init = savi_sd["initializer.slots_mu"] + savi_sd["initializer.slots_sigma"] * noise
        savi.initializer.forward = lambda batch_size, **kw: init   # inject the sampled slots

        # ---- per-stage vectors (frame 0) ----
        feats = savi.encode(videos[:, 0])
        out["encode_feats_sub"] = feats[:, ::m["feat_stride"]].clone()
        s = init
        # SlotAttention per iteration: run with num_iters_first = 1, 2, 3
        for it in (1, 2, 3):
            savi.slot_attention.num_iters_first = it
            out[f"sa_iter{it}"] = savi.slot_attention(inputs=feats, slots=init, step=0).clone()
        savi.slot_attention.num_iters_first = 3
        slots0 = out["sa_iter3"]
        out["sa_step1"] = savi.slot_attention(inputs=feats, slots=init, step=1).clone()
        out["transition"] = savi.transition_module(slots0).clone()

        # ---- evaluator composition (05_evaluate_predictor.py:82-96) ----
        sh = savi(mode="decomp", x=videos, num_imgs=20, decode=False)["slot_history"]
        out["slot_history"] = sh.clone()
        # one predictor step on a full 10-frame window, teacher-forced from encoded slots
        win = sh[:, :10]
        out["pred_step_n10"] = pred.predictor(slots=win, time_step=0, text_embeddings=text).clone()
        out["pred_step_n1"] = pred.predictor(slots=sh[:, :1], time_step=0, text_embeddings=text).clone()
        ps = pred(sh, caption_tokens=None)
        out["pred_slots"] = ps.clone()
        dec = savi(mode="decode", slots=ps.reshape(B * 19, 8, 128))
        out["pred_imgs"] = dec["recons_imgs"].view(B, 19, 3, 64, 64).clamp(0, 1).clone()
        # decoder detail for one frame (first sequence, last step)
        one = savi(mode="decode", slots=ps[:1, -1])
        out["dec_recons"] = one["recons"].clone()
        out["dec_masks"] = one["masks"].clone()
        out["dec_img"] = one["recons_imgs"].clone()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "cater_b2.pt")
    torch.save(out, path)
    tot = sum(v.numel() * 4 for v in out.values() if torch.is_tensor(v))
    print(f"wrote {path}: {tot/1e6:.2f} MB; slot std t0={sh[:,0].std():.3f} pred t18={ps[:,-1].std():.3f}")


if __name__ == "__main__":
    main()
