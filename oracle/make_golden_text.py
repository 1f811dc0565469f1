"""
Golden vectors for the text encoder of TextOCVP_CustomTF from the REAL reference module
(src/models/EncodersDecoders/text_encoders.py, imported read-only).  TEST INFRASTRUCTURE -- build container only:

    python -m oracle.make_golden_text
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_import  # noqa: E402
from textocvp_b200 import weights  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
META = dict(B=5, L=24, seed=18, cap_seed=4, bias_scale=0.02, ln_jitter=0.05)


def main():
    m = META
    _, pred = ref_import.build_reference(num_preds=3)
    enc = pred.predictor.text_encoder.eval()
    sd = weights.text_encoder_state_dict(m["seed"], bias_scale=m["bias_scale"], ln_jitter=m["ln_jitter"])
    enc.load_state_dict(sd, strict=True)
    tokens, lengths = weights.synthetic_captions(m["B"], m["L"], seed=m["cap_seed"])
    with torch.no_grad():
        out = enc(text=tokens, text_length=lengths)
    path = os.path.join(OUT, "text_encoder_b5.pt")
    torch.save({"meta": m, "out": out.clone()}, path)
    print(f"wrote {path}: out {tuple(out.shape)} std {out.std():.3f}")


if __name__ == "__main__":
    main()
