"""Kernel timeline (CUPTI via torch.profiler) of one decode call: start / duration / gap per kernel for two chunks, and the
totals of busy time vs wall time.  Dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from textocvp_b200 import rollout
dev = torch.device("cuda:0")
stage = sys.argv[1] if len(sys.argv) > 1 else "decode"
B = 256
savi, pred, _ = rollout.build_models(dev)
ps = torch.randn(B * 19, 8, 128, device=dev)
sh = torch.randn(B, 20, 8, 128, device=dev)
text = torch.randn(B, 32, 512, device=dev)
fn = (lambda: savi.decode(ps, only_imgs=True)) if stage == "decode" else (lambda: pred(sh, text_embeddings=text))
for _ in range(3): fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in evs)
wall = max(e.time_range.end for e in evs) - t0
print(f"{stage}: {len(evs)} kernels, wall {wall/1e3:.2f} ms, sum of kernel durations {busy/1e3:.2f} ms")
lo, hi = (len(evs) // 2, len(evs) // 2 + 16) if stage == "decode" else (len(evs) - 100, len(evs) - 70)
prev_end = None
for e in evs[lo:hi]:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = (e.time_range.start - prev_end) if prev_end is not None else 0
    print(f"  t={s/1e3:9.3f} ms  dur={d:8.1f} us  gap={gap:7.1f} us  {e.name[:60]}")
    prev_end = max(prev_end or 0, e.time_range.end)
# total idle gaps (no kernel running)
idle, cur_end = 0.0, evs[0].time_range.end
for e in evs[1:]:
    if e.time_range.start > cur_end: idle += e.time_range.start - cur_end
    cur_end = max(cur_end, e.time_range.end)
print(f"idle (no kernel resident) {idle/1e3:.3f} ms")
