"""A/B of the corrector's per-slot update kernel in one process (tocvp_tuning.corrector_mode): 0 = 3xTF32 mma.sync with the
weights streamed through a shared-memory ring (default), 2 = first tensor-core version (weights from L2 in the MMA loop),
1 = fp32 SIMT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights, _lib as L
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev)
B = 256
feats = torch.randn(B, 4096, 128, device=dev).half()
ssd = weights.savi_state_dict(14)
_, _, noise = weights.synthetic_inputs(B, 2, 32, seed=0)
init = (ssd["initializer.slots_mu"] + ssd["initializer.slots_sigma"] * noise).to(dev)
o = torch.empty_like(init); nx = torch.empty_like(init)
def t(iters, nxt, n=10):
    f = lambda: savi.slot_attention.run(feats, 4096 * 128, B, 4096, init, iters, o, 8 * 128, nxt)
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
outs = {}
for rep in range(2):
    for mode in (1, 2, 0):
        setattr(L.TUNING, "corrector_mode", int(mode))
        a, b = t(3, None), t(1, nx)
        outs[mode] = (o.clone(), nx.clone())
        print(f"corrector update mode {mode}: 3 iterations {a*1e3:.0f} us, 1 iteration + transition {b*1e3:.0f} us", flush=True)
setattr(L.TUNING, "corrector_mode", int(0))
d0 = float((outs[0][0] - outs[1][0]).norm() / outs[1][0].norm()); d1 = float((outs[0][1] - outs[1][1]).norm() / outs[1][1].norm())
print(f"relative difference mode 0 vs fp32 SIMT: slots {d0:.2e}, transition {d1:.2e}")
d2 = float((outs[2][0] - outs[1][0]).norm() / outs[1][0].norm())
print(f"relative difference mode 2 vs fp32 SIMT: slots {d2:.2e}")
