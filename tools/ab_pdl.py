"""A/B: predictor rollout with / without programmatic dependent launch (tocvp_tuning.no_pdl), same process, graph replay;
also checks that both give bit-identical predictions."""
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
for rep in range(3):
    for on in (0, 1):
        setattr(L.TUNING, "no_pdl", int(not (on)))
        object.__setattr__(pred.predictor, "_graph", None)      # re-capture with the new launch attribute
        ms, o = t()
        outs[on] = o
        print(f"predict PDL={'on' if on else 'off'}: {ms:.2f} ms", flush=True)
print("bit-identical:", torch.equal(outs[0], outs[1]), flush=True)
for on in (0, 1):                                                # eager (no graph) as well
    setattr(L.TUNING, "no_pdl", int(not (on)))
    pred.predictor.use_cuda_graph = False
    ms, o = t()
    print(f"eager predict PDL={'on' if on else 'off'}: {ms:.2f} ms; equal to graph: {torch.equal(o, outs[1])}", flush=True)
setattr(L.TUNING, "no_pdl", int(not (1)))
