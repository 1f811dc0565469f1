import sys; sys.path.insert(0, "/root/repo")
import torch
from textocvp_b200 import rollout
raw = torch.rand(256, 19, 3, 64, 64, device="cuda"); vid = torch.rand(256, 20, 3, 64, 64, device="cuda")
for want in (True, False):
    for _ in range(3): rollout.frame_metrics(raw, vid, 1, clamp=True, want_ssim=want)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): rollout.frame_metrics(raw, vid, 1, clamp=True, want_ssim=want)
    e1.record(); torch.cuda.synchronize()
    print("want_ssim", want, e0.elapsed_time(e1) / 10, "ms")
