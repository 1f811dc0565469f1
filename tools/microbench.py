"""Per-kernel micro-benchmarks (CUDA events, warm-up, L2-larger-than-cache inputs). Dev tool."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import ops


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_gemm():
    for M, N, K in [(20480, 1536, 512), (20480, 512, 512), (20480, 2048, 512), (20480, 512, 2048), (2048, 1536, 512),
                    (8192, 8192, 8192)]:
        a = torch.randn(M, K, device="cuda").half()
        w = torch.randn(N, K, device="cuda").half()
        ms = timeit(lambda: ops.gemm_f16(a, w, out_f32=False, out_f16=True))
        ms_t = timeit(lambda: a @ w.t())
        print(f"gemm {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s   (torch/cuBLAS {2*M*N*K/ms_t/1e9:.1f})")


def bench_conv():
    for n, c in [(2048, 64), (2048, 32)]:
        x = torch.randn(n, 64, 64, c, device="cuda").half()
        w = ops.pack_conv5x5_weight(torch.randn(c, c, 5, 5, device="cuda") * 0.02)
        b = torch.zeros(c, device="cuda")
        ms = timeit(lambda: ops.conv5x5_f16(x, w, b))
        fl = 2 * n * 4096 * 25 * c * c
        print(f"conv5x5 {c}->{c} n={n}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "conv"]
    for w in which:
        globals()["bench_" + w]()
