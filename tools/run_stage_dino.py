"""Run one CLIPort-shape stage repeatedly (for ncu launch lists). usage: run_stage_dino.py {predict|decode|decomp} [B] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights
stage = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 128; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
NP = 29
dev = torch.device("cuda:0")
dino, pred, _ = rollout.build_dino_models(dev, num_preds=NP)
feats, text, noise = weights.synthetic_dino_inputs(B, NP + 1, 81, L=16, seed=0)
feats, text = feats.to(dev), text.to(dev)
sh = torch.randn(B, NP + 1, 10, 128, device=dev)
ps = torch.randn(B * NP, 10, 128, device=dev)
init = torch.randn(B, 10, 128, device=dev)
for _ in range(reps):
    if stage == "predict": pred(sh, text_embeddings=text)
    elif stage == "decode": dino.decode(ps, only_imgs=True)
    elif stage == "decomp": dino(mode="decomp", x=feats, num_imgs=NP + 1, decode=False, init_slots=init)
torch.cuda.synchronize()
print("done")
