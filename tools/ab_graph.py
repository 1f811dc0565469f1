"""A/B: predictor rollout enqueued eagerly vs replayed from a CUDA graph (same process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev)
_, text, _ = weights.synthetic_inputs(256, 20, 32, seed=0)
text = text.to(dev)
sh = torch.randn(256, 20, 8, 128, device=dev)
def t(n=3):
    pred(sh, text_embeddings=text); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): o = pred(sh, text_embeddings=text)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, o
outs = {}
for rep in range(2):
    for g in (False, True):
        pred.predictor.use_cuda_graph = g
        ms, o = t()
        outs[g] = o
        print(f"predict cuda_graph={g}: {ms:.2f} ms", flush=True)
print("max abs diff eager vs graph:", float((outs[True] - outs[False]).abs().max()))
