"""A/B of the decoder in one process (tocvp_tuning.decode_mode bits): 0 = default, 1 = layer 1 generated inside the layer-2
conv, 8 = serial chunks (no layer-1 / conv overlap), 2 = first-version head conv3x3 (shifted windows, N = 16) instead of the taps-in-N kernel;
per-layer conv times from the event pairs tocvp_savi_decode records."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, _lib as L
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev)
ps = torch.randn(256 * 19, 8, 128, device=dev)
n_chunks = 19
def t(n=3):
    savi.decode(ps, only_imgs=True); torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * 3 * n_chunks)]
    for e in evs: e.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): savi.decode(ps, only_imgs=True, conv_events=evs)
    e1.record(); torch.cuda.synchronize()
    per = [a.elapsed_time(b) for a, b in zip(evs[0::2], evs[1::2])]
    lay = [sum(per[l::3]) / n_chunks for l in range(3)]
    return e0.elapsed_time(e1) / n, lay
for rep in range(2):
    for mode in (0, 8, 0, 8):
        setattr(L.TUNING, "decode_mode", int(mode))
        ms, lay = t()
        print(f"decode mode {mode}: {ms:.1f} ms; conv layers 2/3/4: {lay[0]:.3f} {lay[1]:.3f} {lay[2]:.3f} ms", flush=True)
setattr(L.TUNING, "decode_mode", int(0))
