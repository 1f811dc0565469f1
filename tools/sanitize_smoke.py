"""Tiny end-to-end pass of both model families (compute-sanitizer target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev, num_context=1, num_preds=2)
pred.predictor.use_cuda_graph = False
videos, text, noise = weights.synthetic_inputs(2, 3, 32, seed=1)
out = rollout.forward_eval(savi, pred, videos.to(dev), text.to(dev), 1, 2)
dino, dpred, _ = rollout.build_dino_models(dev, num_preds=2)
dpred.predictor.use_cuda_graph = False
feats, dtext, _ = weights.synthetic_dino_inputs(2, 3, 81, L=16, seed=0)
o2 = rollout.forward_eval_dino(dino, dpred, feats.to(dev), dtext.to(dev), 1, 2)
# multi-chunk decode: the chunk-pipelined driver (side stream, three activation buffers) with a ragged tail
d2 = savi.decode(torch.randn(256 + 8, 8, 128, device=dev), only_imgs=True)
torch.cuda.synchronize()
print("ok", float(out["psnr"].mean()), tuple(o2["pred_imgs"].shape))
