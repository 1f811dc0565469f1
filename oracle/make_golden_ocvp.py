"""
Golden vectors for the sibling predictors (VanillaTransformer, OCVPSeq, OCVPPar) from the REAL reference modules
(src/models/Predictors/OCVP.py through lib/setup_model.setup_predictor, imported read-only; OCVPPar has no factory entry
in the reference and is instantiated directly with OCVPSeq.json's parameters inside the reference's PredictorWrapper).
TEST INFRASTRUCTURE -- build container only:    python -m oracle.make_golden_ocvp
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
META = dict(B=3, n=7, S=8, seed=19, input_seed=6, bias_scale=0.02, ln_jitter=0.05, num_context=2, num_preds=4,
            input_buffer_size=10)


def main():
    m = META
    ref_import._prepare()
    cwd = os.getcwd()
    os.chdir(ref_import.REF_ROOT)
    out = {"meta": m}
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import lib.setup_model as sm
            from CONFIG import DEFAULTS
        g = torch.Generator().manual_seed(m["input_seed"])
        slots = torch.randn(m["B"], m["n"], m["S"], 128, generator=g)
        hist = torch.randn(m["B"], m["num_context"] + m["num_preds"], m["S"], 128, generator=g)
        out["slots"], out["hist"] = slots, hist
        for kind in ("VanillaTransformer", "OCVPSeq", "OCVPPar"):
            with contextlib.redirect_stdout(io.StringIO()):
                cfg = "OCVPSeq" if kind == "OCVPPar" else kind
                exp = {"model": {"model_name": "SAVi", "model_params": json.load(open("src/configs/models/SAVi.json"))},
                       "predictor": json.load(open(f"src/configs/predictors/{cfg}.json")),
                       "prediction_params": {**DEFAULTS["prediction_params"], "num_context": m["num_context"],
                                             "num_preds": m["num_preds"], "input_buffer_size": m["input_buffer_size"]}}
                if kind == "OCVPPar":
                    from models.Predictors.OCVP import OCVPPar
                    from models.Predictors.predictor_wrapper import PredictorWrapper
                    pp = exp["predictor"]["predictor_params"]
                    body = OCVPPar(num_slots=exp["model"]["model_params"]["num_slots"],
                                   slot_dim=exp["model"]["model_params"]["slot_dim"], token_dim=pp["token_dim"],
                                   hidden_dim=pp["hidden_dim"], num_layers=pp["num_layers"], n_heads=pp["n_heads"],
                                   residual=pp["residual"], input_buffer_size=m["input_buffer_size"])
                    pred = PredictorWrapper(exp_params=copy.deepcopy(exp), predictor=body).eval()
                else:
                    pred = sm.setup_predictor(copy.deepcopy(exp)).eval()
            sd = weights.ocvp_state_dict(kind, m["seed"], bias_scale=m["bias_scale"], ln_jitter=m["ln_jitter"])
            pred.predictor.load_state_dict(sd, strict=True)
            pred.encode_text_caption = lambda **kw: None
            with torch.no_grad():
                out[kind + "_step"] = pred.predictor(slots=slots).clone()
                out[kind + "_rollout"] = pred(hist).clone()
    finally:
        os.chdir(cwd)
    path = os.path.join(OUT, "ocvp_b3.pt")
    torch.save(out, path)
    print(f"wrote {path}:", {k: tuple(v.shape) for k, v in out.items() if torch.is_tensor(v)})


if __name__ == "__main__":
    main()
