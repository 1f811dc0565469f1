"""A/B: predictor rollout with / without the W-resident GEMM variant (same process, graph replay)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights, ops
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev)
_, text, _ = weights.synthetic_inputs(256, 20, 32, seed=0)
text = text.to(dev)
sh = torch.randn(256, 20, 8, 128, device=dev)
def t(n=5):
    pred(sh, text_embeddings=text); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): o = pred(sh, text_embeddings=text)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(3):
    for mode in (258, 259):
        ops.set_gemm_mode(mode)
        object.__setattr__(pred.predictor, "_graph", None)      # re-capture with the new kernel choice
        print(f"predict W-resident={'on' if mode == 259 else 'off'}: {t():.2f} ms", flush=True)
ops.set_gemm_mode(259)
