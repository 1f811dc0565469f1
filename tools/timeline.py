"""Kernel timeline (CUPTI via torch.profiler) of one call of a stage (decode | predict | decomp | eval): start / duration / gap per kernel for a window, the
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
videos = torch.rand(B, 20, 3, 64, 64, device=dev)
init = torch.randn(B, 8, 128, device=dev)
fn = {"decode": lambda: savi.decode(ps, only_imgs=True),
      "predict": lambda: pred(sh, text_embeddings=text),
      "decomp": lambda: savi(mode="decomp", x=videos, num_imgs=20, decode=False, init_slots=init),
      "eval": lambda: rollout.forward_eval(savi, pred, videos, text, 1, 19, init_slots=init)}[stage]
if len(sys.argv) > 2:                      # optional: ops.set_gemm_mode(<mode>) before the run (e.g. 260 / 261)
    from textocvp_b200 import ops
    ops.set_gemm_mode(int(sys.argv[2]))
if os.environ.get("TOCVP_TUNING"):          # e.g. TOCVP_TUNING=decode_mode=16,no_pdl=1
    from textocvp_b200 import ops
    ops.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in os.environ["TOCVP_TUNING"].split(","))})
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
lo, hi = {"decode": (len(evs) // 2, len(evs) // 2 + 16), "predict": (len(evs) - 100, len(evs) - 70),
          "decomp": (0, 40), "eval": (0, 0)}[stage]
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
# time per kernel name (durations include the griddepcontrol.wait overlap of programmatic dependent launches)
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    agg[e.name[:78]][0] += 1
    agg[e.name[:78]][1] += e.time_range.end - e.time_range.start
for name, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"  {tot/1e3:8.2f} ms  x{cnt:5d}  avg {tot/cnt:7.1f} us  {name}")
# the largest idle gaps and what follows them
gaps, cur_end = [], evs[0].time_range.end
for e in evs[1:]:
    if e.time_range.start > cur_end: gaps.append((e.time_range.start - cur_end, e.time_range.start - t0, e.name[:50]))
    cur_end = max(cur_end, e.time_range.end)
for g, at, name in sorted(gaps, reverse=True)[:8]:
    print(f"  gap {g:8.1f} us at t={at/1e3:8.3f} ms before {name}")
