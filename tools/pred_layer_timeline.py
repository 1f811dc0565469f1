"""Per-kernel durations of the LAST full-window predictor step (CUPTI): one line per kernel of layers 3-4, plus per-name totals
of the whole rollout.  Dev tool."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from textocvp_b200 import rollout
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
savi, pred, _ = rollout.build_models(dev)
sh = torch.randn(B, 20, 8, 128, device=dev)
text = torch.randn(B, 32, 512, device=dev)
if os.environ.get("TOCVP_TUNING"):
    from textocvp_b200 import ops
    ops.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in os.environ["TOCVP_TUNING"].split(","))})
for _ in range(3): pred(sh, text_embeddings=text)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pred(sh, text_embeddings=text); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
wall = max(e.time_range.end for e in evs) - t0
print(f"predict: {len(evs)} kernels, wall {wall/1e3:.2f} ms")
# last step = the kernels after the second-to-last commit_prediction_kernel
idx = [i for i, e in enumerate(evs) if "commit_prediction" in e.name]
lo = idx[-2] + 1
step = evs[lo:idx[-1] + 1]
print(f"last step: {len(step)} kernels, {(step[-1].time_range.end - step[0].time_range.start)/1e3:.3f} ms")
prev_end = None
for e in step[2 + 12 * 3: 2 + 12 * 5]:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = (e.time_range.start - prev_end) if prev_end is not None else 0
    print(f"  dur={d:7.1f} us  gap={gap:6.1f} us  {e.name[:70]}")
    prev_end = max(prev_end or 0, e.time_range.end)
agg = collections.defaultdict(lambda: [0, 0.0])
for e in step:
    agg[e.name[:60]][0] += 1
    agg[e.name[:60]][1] += e.time_range.end - e.time_range.start
for name, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  last step: {tot:8.1f} us  x{cnt:4d}  avg {tot/cnt:7.1f} us  {name}")
