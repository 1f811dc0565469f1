import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ms = t(lambda: x.zero_()); print(f"memset 1 GiB: {ms:.3f} ms -> {1.0737/ms*1e3:.0f} GB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy 1 GiB: {ms:.3f} ms -> {2*1.0737/ms*1e3:.0f} GB/s read+write")
ms = t(lambda: x.sum()); print(f"read-reduce 1 GiB (uint8 sum): {ms:.3f} ms")
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print(f"read-reduce 1 GiB (f32 sum): {ms:.3f} ms -> {1.0737/ms*1e3:.0f} GB/s read")
