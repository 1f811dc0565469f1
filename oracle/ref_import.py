"""
Import recipe for the REAL reference (read-only, /root/reference) -- build container only.

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` to produce the golden vectors under
``tests/golden/`` and by ``tests/test_oracle_vs_reference.py`` (skipped when /root/reference
is absent, as on the GPU box).  Nothing is copied from the reference: it is imported in place
with stub modules for its missing third-party dependencies (timm, nltk), following
SURVEY.md Appendix C.
"""
from __future__ import annotations

import contextlib
import copy
import io
import json
import os
import sys
import types

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "models"))


_done = False


def _prepare():
    global _done
    if _done:
        return
    from transformers import T5EncoderModel  # noqa: F401  (must precede the timm stub)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("nltk", download=lambda *a, **k: False, word_tokenize=str.split)
    stub("timm", create_model=None)
    stub("timm.models", layers=types.SimpleNamespace(GroupNorm=None),
         resnet=types.SimpleNamespace(BasicBlock=None))
    stub("timm.models.vision_transformer", _create_vision_transformer=None,
         VisionTransformer=type("VT", (), {}))
    sys.path.insert(0, os.path.join(REF_ROOT, "src"))
    _done = True


def build_reference(num_preds: int = 19, num_context: int = 1, savi_overrides=None, pred_overrides=None):
    """Instantiate the reference SAVi and PredictorWrapper(TextOCVP_CustomTF) from its own JSONs."""
    _prepare()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)          # CONFIG.py paths are cwd-relative
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import lib.setup_model as sm
            from CONFIG import DEFAULTS
            mp = json.load(open("src/configs/models/SAVi.json"))
            if savi_overrides:
                mp = _deep_update(mp, savi_overrides)
            pp = json.load(open("src/configs/predictors/TextOCVP_CustomTF.json"))
            if pred_overrides:
                pp = _deep_update(pp, pred_overrides)
            exp = {"model": {"model_name": "SAVi", "model_params": mp},
                   "predictor": pp,
                   "prediction_params": {**DEFAULTS["prediction_params"],
                                         "num_context": num_context, "num_preds": num_preds}}
            savi = sm.setup_model(copy.deepcopy(exp["model"])).eval()
            pred = sm.setup_predictor(copy.deepcopy(exp)).eval()
    finally:
        os.chdir(cwd)
    return savi, pred


def build_reference_dino(img_size: int = 128, num_patches: int = 81, num_preds: int = 29, num_context: int = 1):
    """Reference ExtendedDINOSAUR (frozen ViT backbone replaced by identity: it is fed patch features directly, SURVEY.md
    Appendix C) and PredictorWrapper(TextOCVP_CustomTF), from the reference's own JSONs with the BASELINE.json geometry."""
    _prepare()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import torch
            import lib.setup_model as sm
            import models.ExtendedDINOSAUR as ED
            from CONFIG import DEFAULTS
            ED.get_encoder = lambda **kw: torch.nn.Identity()
            mp = json.load(open("src/configs/models/ExtendedDINOSAUR.json"))
            mp["img_size"] = img_size
            mp["decoder"]["decoder_params"]["num_patches"] = num_patches
            pp = json.load(open("src/configs/predictors/TextOCVP_CustomTF.json"))
            exp = {"model": {"model_name": "ExtendedDINOSAUR", "model_params": mp},
                   "predictor": pp,
                   "prediction_params": {**DEFAULTS["prediction_params"],
                                         "num_context": num_context, "num_preds": num_preds}}
            dino = sm.setup_model(copy.deepcopy(exp["model"])).eval()
            pred = sm.setup_predictor(copy.deepcopy(exp)).eval()
    finally:
        os.chdir(cwd)
    return dino, pred


def _deep_update(d, u):
    d = copy.deepcopy(d)
    for k, v in u.items():
        if isinstance(v, dict) and isinstance(d.get(k), dict):
            d[k] = _deep_update(d[k], v)
        else:
            d[k] = v
    return d


def load_weights(savi, pred, savi_sd, pred_sd):
    """Strict load of our generated dicts into the reference modules (proves the key contract)."""
    if savi is not None:
        savi.load_state_dict(savi_sd, strict=True)
    body = pred.predictor
    full = dict(body.state_dict())
    for k, v in pred_sd.items():
        if k not in full:
            raise KeyError(f"predictor key {k} not in reference state_dict")
        full[k] = v
    missing = [k for k in full if k not in pred_sd and not k.startswith("text_encoder.")]
    if missing:
        raise KeyError(f"reference predictor keys not generated: {missing}")
    body.load_state_dict(full, strict=True)
