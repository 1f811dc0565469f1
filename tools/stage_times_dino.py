"""CUDA-event timing of the CLIPort / ExtendedDINOSAUR rollout stages (BASELINE.json configs[3]; dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 29
if len(sys.argv) > 3:                     # optional: PDL on (1) or off (0)
    from textocvp_b200 import _lib as L
    setattr(L.TUNING, "no_pdl", int(not (int(sys.argv[3]))))
dev = torch.device("cuda:0")
dino, pred, _ = rollout.build_dino_models(dev, num_preds=NP)
feats, text, noise = weights.synthetic_dino_inputs(B, NP + 1, 81, L=16, seed=0)
feats, text = feats.to(dev), text.to(dev)
dsd = weights.dino_state_dict(16)
init = (dsd["initializer.slots_mu"] + dsd["initializer.slots_sigma"] * noise).to(dev)


def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r

ms_dec, out = t(lambda: dino(mode="decomp", x=feats, num_imgs=NP + 1, decode=False, init_slots=init))
sh = out["slot_history"]
ms_proj, _ = t(lambda: dino.project(feats))
ms_pred, ps = t(lambda: pred(sh, text_embeddings=text))
ms_decode, _ = t(lambda: dino.decode(ps.reshape(B * NP, 10, 128), only_imgs=True))
ms_all, _ = t(lambda: rollout.forward_eval_dino(dino, pred, feats, text, 1, NP, init_slots=init))
fl = B * NP * (4.89e9 + 10.46e9) + B * 234.7e9 * NP / 29
print(f"CLIPort B={B} T={NP+1}: decomp {ms_dec:.1f} ms [project {ms_proj:.1f}] | predict {ms_pred:.1f} ms | "
      f"decode {ms_decode:.1f} ms | step {ms_all:.1f} ms -> {B*NP/ms_all*1e3:.0f} frames/s "
      f"({fl/ms_all/1e9:.0f} TFLOP/s dense-formulation)")
