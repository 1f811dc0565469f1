"""A/B of the decoder in one process (tocvp_tuning.decode_mode bits): 0 = default (compositing fused into the head conv),
16 = separate compositing kernel, 8 = serial chunks, 1 = layer 1 generated inside the layer-2 conv, 2 / 4 = first versions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, _lib as L
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 16, 0, 16]
savi, pred, _ = rollout.build_models(dev)
slots = torch.randn(B * 19, 8, 128, device=dev)
def t(only, n=3):
    savi.decode(slots, only_imgs=only); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): savi.decode(slots, only_imgs=only)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mode in modes:
    setattr(L.TUNING, "decode_mode", int(mode))
    print(f"decode_mode={mode}: full outputs {t(False):.2f} ms, images only {t(True):.2f} ms", flush=True)
setattr(L.TUNING, "decode_mode", 0)
