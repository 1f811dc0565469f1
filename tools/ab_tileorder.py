"""A/B: predictor rollout with / without alternating tile traversal (tocvp_tuning.no_tile_alternation), same process, graph replay;
checks bit-identical predictions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights, _lib as L
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
savi, pred, _ = rollout.build_models(dev)
_, text, _ = weights.synthetic_inputs(B, 20, 32, seed=0)
text = text.to(dev)
sh = torch.randn(B, 20, 8, 128, device=dev)
def t(n=5):
    pred(sh, text_embeddings=text); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): o = pred(sh, text_embeddings=text)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, o.clone()
outs = {}
for rep in range(4):
    for on in (0, 1):
        setattr(L.TUNING, "no_tile_alternation", int(not (on)))
        object.__setattr__(pred.predictor, "_graph", None)      # re-capture
        ms, o = t()
        outs[on] = o
        print(f"predict alternating tile order={'on' if on else 'off'}: {ms:.2f} ms", flush=True)
print("bit-identical:", torch.equal(outs[0], outs[1]), flush=True)
setattr(L.TUNING, "no_tile_alternation", int(not (1)))
