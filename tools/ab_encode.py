"""A/B of the encoder variants in one process (tocvp_tuning.encode_mode bits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights, _lib as L
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev)
videos, _, _ = weights.synthetic_inputs(256, 20, 32, seed=0)
xs = videos.to(dev).reshape(256 * 20, 3, 64, 64)
def t(n=3):
    savi._encode_raw(xs, 5120, 3 * 64 * 64, False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): savi._encode_raw(xs, 5120, 3 * 64 * 64, False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(2):
    for mode in (0, 8, 4, 0, 8):
        setattr(L.TUNING, "encode_mode", int(mode))
        print(f"encode mode {mode}: {t():.2f} ms", flush=True)
setattr(L.TUNING, "encode_mode", int(0))
