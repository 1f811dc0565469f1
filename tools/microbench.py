"""Per-kernel micro-benchmarks (CUDA events, warm-up, L2-larger-than-cache inputs). Dev tool."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import ops


_big = None


def timeit(fn, iters=20, warm=3):
    """Device time per call.  A ~10 ms filler kernel is queued first so that the host has enqueued all `iters` launches
    before the GPU reaches them: short kernels are then timed back to back, not at the host's launch rate."""
    global _big
    if _big is None:
        _big = torch.randn(16384, 16384, device="cuda").half()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _big @ _big
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_gemm():
    """single-CTA (mode 1) vs CTA-pair kernel with 128 / 256-wide tiles vs the automatic choice (mode 0)"""
    for M, N, K in [(20480, 1536, 512), (20480, 512, 512), (20480, 2048, 512), (20480, 512, 2048), (10240, 2048, 512),
                    (4096, 1536, 512), (207360, 1024, 1024), (8192, 8192, 8192)]:
        a = torch.randn(M, K, device="cuda").half()
        w = torch.randn(N, K, device="cuda").half()
        res = []
        for mode in (1, 128, 256, 0):
            ops.set_gemm_mode(mode)
            ms = timeit(lambda: ops.gemm_f16(a, w, out_f32=False, out_f16=True))
            res.append(f"mode{mode}: {ms*1e3:7.1f} us {2*M*N*K/ms/1e9:7.1f} TF")
        ops.set_gemm_mode(256)                                   # 256-wide tiles without the W-resident variant
        ops.set_tuning(gemm_no_wres=1)
        ms = timeit(lambda: ops.gemm_f16(a, w, out_f32=False, out_f16=True))
        res.append(f"256/noWres: {ms*1e3:7.1f} us {2*M*N*K/ms/1e9:7.1f} TF")
        ops.set_tuning(gemm_no_wres=0)
        ops.set_gemm_mode(0)
        ms_t = timeit(lambda: a @ w.t())
        print(f"gemm {M}x{N}x{K}: " + " | ".join(res) + f" | cuBLAS {2*M*N*K/ms_t/1e9:.1f}", flush=True)


def bench_conv():
    for n, c in [(2048, 64), (2048, 32)]:
        x = torch.randn(n, 64, 64, c, device="cuda").half()
        w = ops.pack_conv5x5_weight(torch.randn(c, c, 5, 5, device="cuda") * 0.02)
        b = torch.zeros(c, device="cuda")
        fl = 2 * n * 4096 * 25 * c * c
        for mode in (1, 0):
            ops.set_conv_mode(mode)
            ms = timeit(lambda: ops.conv5x5_f16(x, w, b))
            print(f"conv5x5 {c}->{c} n={n} mode{mode}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
        ops.set_conv_mode(0)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "conv"]
    for w in which:
        globals()["bench_" + w]()
