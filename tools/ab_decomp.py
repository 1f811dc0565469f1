"""A/B of the 20-frame decomp (encode + corrector chain) in one process: chained corrector (tocvp_slot_attention_seq, two
kernels per frame, one library call per encode chunk) vs one library call per frame (three kernels per frame)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
savi, pred, _ = rollout.build_models(dev)
videos, text, noise = weights.synthetic_inputs(B, 20, 32, seed=0)
videos = videos.to(dev)
ssd = weights.savi_state_dict(14)
init = (ssd["initializer.slots_mu"] + ssd["initializer.slots_sigma"] * noise).to(dev)
def t(n=5):
    savi(mode="decomp", x=videos, num_imgs=20, decode=False, init_slots=init); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): o = savi(mode="decomp", x=videos, num_imgs=20, decode=False, init_slots=init)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, o["slot_history"].clone()
outs = {}
for rep in range(3):
    for chain in (False, True):
        savi.chain_corrector = chain
        ms, o = t()
        outs[chain] = o
        print(f"decomp (20 frames, B={B}) chained={chain}: {ms:.2f} ms", flush=True)
print("bit-identical:", torch.equal(outs[False], outs[True]))
