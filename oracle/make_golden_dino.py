"""
Golden vectors for the CLIPort / ExtendedDINOSAUR shape (BASELINE.json configs[3]) from the REAL reference modules
(imported read-only from /root/reference; the frozen ViT backbone is replaced by identity and fed synthetic patch
features, per the north star).  TEST INFRASTRUCTURE -- build container only:

    python -m oracle.make_golden_dino

Weights are regenerated from seeds (textocvp_b200.weights.dino_state_dict) and loaded STRICTLY into the reference.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_import  # noqa: E402
from textocvp_b200 import weights  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

META = dict(B=2, T=4, N=81, L=16, img_size=128, dino_seed=16, pred_seed=17, input_seed=3, bias_scale=0.02,
            ln_jitter=0.05, bn_jitter=0.2, mlp_out_scale=0.1, num_preds=3, num_context=1)


def main():
    torch.set_num_threads(os.cpu_count())
    m = META
    dsd = weights.dino_state_dict(m["dino_seed"], img_size=m["img_size"], num_patches=m["N"],
                                  bias_scale=m["bias_scale"], ln_jitter=m["ln_jitter"], bn_jitter=m["bn_jitter"])
    psd = weights.predictor_state_dict(m["pred_seed"], mlp_out_scale=m["mlp_out_scale"], ln_jitter=m["ln_jitter"])
    feats, text, noise = weights.synthetic_dino_inputs(m["B"], m["T"], m["N"], L=m["L"], seed=m["input_seed"])
    dino, pred = ref_import.build_reference_dino(m["img_size"], m["N"], m["num_preds"], m["num_context"])
    dino.load_state_dict(dsd, strict=True)
    ref_import.load_weights(None, pred, None, psd)
    pred.encode_text_caption = lambda **kw: text
    B = m["B"]
    out = {"meta": m}
    with torch.no_grad():
        init = dsd["initializer.slots_mu"] + dsd["initializer.slots_sigma"] * noise
        dino.initializer.forward = lambda batch_size, **kw: init
        proj = dino.linear_feat_proj(feats[:, 0])
        out["proj_feats"] = proj.clone()
        out["sa_step0"] = dino.slot_attention(inputs=proj, slots=init, step=0).clone()
        out["sa_step1"] = dino.slot_attention(inputs=proj, slots=init, step=1).clone()
        out["transition"] = dino.transition_module(out["sa_step0"]).clone()
        sh = dino(mode="decomp", x=feats, num_imgs=m["T"], decode=False)["slot_history"]
        out["slot_history"] = sh.clone()
        ps = pred(sh, caption_tokens=None)
        out["pred_slots"] = ps.clone()
        dec = dino(mode="decode", slots=ps.reshape(B * m["num_preds"], 10, 128))
        out["pred_imgs"] = dec["recons_imgs"].clone()
        out["pred_feats"] = dec["recons_feats"].clone()
        out["pred_masks"] = dec["masks"].clone()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "cliport_b2.pt")
    torch.save(out, path)
    tot = sum(v.numel() * 4 for v in out.values() if torch.is_tensor(v))
    print(f"wrote {path}: {tot/1e6:.2f} MB; slot std {sh.std():.3f} img range [{dec['recons_imgs'].min():.2f}, "
          f"{dec['recons_imgs'].max():.2f}] feats std {dec['recons_feats'].std():.3f}")


if __name__ == "__main__":
    main()
