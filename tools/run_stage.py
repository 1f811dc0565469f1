"""Run one stage repeatedly (for ncu launch lists). usage: run_stage.py {predict|decode|encode|decomp} [B] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights
stage = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 256; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev)
videos, text, noise = weights.synthetic_inputs(B, 20, 32, seed=0)
videos, text = videos.to(dev), text.to(dev)
ssd = weights.savi_state_dict(14)
init = (ssd["initializer.slots_mu"] + ssd["initializer.slots_sigma"] * noise).to(dev)
sh = torch.randn(B, 20, 8, 128, device=dev)
ps = torch.randn(B * 19, 8, 128, device=dev)
if os.environ.get("TOCVP_TUNING"):          # e.g. TOCVP_TUNING=decode_mode=16,no_pdl=1
    from textocvp_b200 import ops
    ops.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in os.environ["TOCVP_TUNING"].split(","))})
for _ in range(reps):
    if stage == "predict": pred(sh, text_embeddings=text)
    elif stage == "decode": savi.decode(ps)
    elif stage == "encode": savi._encode_raw(videos.reshape(B * 20, 3, 64, 64), B * 20, 3 * 64 * 64, False)
    elif stage == "decomp": savi(mode="decomp", x=videos, num_imgs=20, decode=False, init_slots=init)
torch.cuda.synchronize()
print("done")
