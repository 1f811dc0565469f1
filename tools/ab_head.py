"""A/B of the decoder head: compositing fused into the head conv (default) vs separate kernel (decode_mode bit 4), kernel
durations from CUPTI in ONE process, alternating; serial chunks (bit 3) so that nothing else shares the HBM with the head."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from textocvp_b200 import rollout, _lib as L
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev)
ps = torch.randn(256 * 19, 8, 128, device=dev)
base = int(sys.argv[1]) if len(sys.argv) > 1 else 8
only = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
for rep in range(2):
    for mode in (base, base | 16):
        L.TUNING.decode_mode = mode
        savi.decode(ps, only_imgs=only); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            savi.decode(ps, only_imgs=only); torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0, 0.0])
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        for e in evs:
            agg[e.name.split("(")[0][-40:]][0] += 1
            agg[e.name.split("(")[0][-40:]][1] += e.time_range.end - e.time_range.start
        wall = max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)
        print(f"decode_mode={mode} only_imgs={only}: wall {wall/1e3:.2f} ms | " +
              " | ".join(f"{n} x{c} avg {t/c:.1f} us" for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:5]), flush=True)
L.TUNING.decode_mode = 0
