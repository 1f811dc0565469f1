"""
Golden vectors for the T5 text-encoder hook from the REAL reference wrapper
(src/models/Predictors/predictor_wrapper.py:90-127 and text_cond_OCVP.py:139-151, imported read-only).
TEST INFRASTRUCTURE -- build container only:

    python -m oracle.make_golden_t5

The reference instantiates ``T5EncoderModel.from_pretrained("t5-small")``, which needs the network.  Here that one call is
redirected to a T5 encoder of the same family (d_model 512) with seeded weights (textocvp_b200.weights.t5_encoder); the
hook under test -- kwargs handling, the encoder call, ``last_hidden_state`` handed to the predictor, the rollout that
consumes it -- is the reference's own code.
"""
from __future__ import annotations

import contextlib
import copy
import io
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_import  # noqa: E402
from textocvp_b200 import weights  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
META = dict(B=2, L=12, t5_seed=19, cap_seed=5, pred_seed=15, mlp_out_scale=0.1, ln_jitter=0.05, num_preds=3, hist_seed=7)


def build_reference_t5(num_preds):
    ref_import._prepare()
    import transformers
    real = transformers.T5EncoderModel.from_pretrained
    transformers.T5EncoderModel.from_pretrained = classmethod(lambda cls, *a, **k: weights.t5_encoder(META["t5_seed"]))
    cwd = os.getcwd()
    os.chdir(ref_import.REF_ROOT)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import lib.setup_model as sm
            from CONFIG import DEFAULTS
            mp = json.load(open("src/configs/models/SAVi.json"))
            pp = json.load(open("src/configs/predictors/TextOCVP_T5.json"))
            exp = {"model": {"model_name": "SAVi", "model_params": mp}, "predictor": pp,
                   "prediction_params": {**DEFAULTS["prediction_params"], "num_context": 1, "num_preds": num_preds}}
            pred = sm.setup_predictor(copy.deepcopy(exp)).eval()
    finally:
        os.chdir(cwd)
        transformers.T5EncoderModel.from_pretrained = real
    return pred


def main():
    m = META
    pred = build_reference_t5(m["num_preds"])
    psd = weights.predictor_state_dict(m["pred_seed"], mlp_out_scale=m["mlp_out_scale"], ln_jitter=m["ln_jitter"])
    full = dict(pred.predictor.state_dict())
    for k, v in psd.items():
        assert k in full, k
        full[k] = v
    pred.predictor.load_state_dict(full, strict=True)
    ids, mask = weights.synthetic_t5_captions(m["B"], m["L"], seed=m["cap_seed"])
    g = torch.Generator().manual_seed(m["hist_seed"])
    slot_history = torch.randn(m["B"], 1 + m["num_preds"], 8, 128, generator=g)
    with torch.no_grad():
        text = pred.encode_text_caption(caption_tokens=ids, attn_masks=mask)
        preds = pred(slot_history, caption_tokens=ids, attn_masks=mask, caption=["a"] * m["B"])
    keys = sorted(k for k in pred.predictor.state_dict() if k.startswith("text_encoder."))
    path = os.path.join(OUT, "t5_hook_b2.pt")
    torch.save({"meta": m, "text_embeddings": text.clone(), "pred_slots": preds.clone(), "text_encoder_keys": keys}, path)
    print(f"wrote {path}: text {tuple(text.shape)} std {text.std():.3f}, preds {tuple(preds.shape)} std {preds.std():.3f}")


if __name__ == "__main__":
    main()
