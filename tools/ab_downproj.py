"""Down-projection shape of the predictor (M = 20480, N = 512, K = 2048) in isolation: tile width 128 vs 256, f16-only output
vs the fp32 residual-in / fp32-out producer form, A rotated over buffers larger than L2.  Dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import ops
M, N, K = 20480, 512, 2048
if len(sys.argv) > 3: M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
As = [torch.randn(M, K, device="cuda").half() for _ in range(4)]
w = (torch.randn(N, K, device="cuda") / 45).half()
b = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda")

def run(label, mode, **kw):
    ops.set_gemm_mode(mode)
    for i in range(4): ops.gemm_f16(As[i % 4], w, bias=b, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record()
    for i in range(n): ops.gemm_f16(As[i % 4], w, bias=b, **kw)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    print(f"{label:40s} mode {mode:3d}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")

if os.environ.get("TOCVP_TUNING"):
    ops.set_tuning(**{k: int(v) for k, v in (kv.split("=") for kv in os.environ["TOCVP_TUNING"].split(","))})
# correctness of the producer form at both widths against fp32 torch
ref = As[0].float() @ w.float().t() + b + res
for mode in (128, 256):
    ops.set_gemm_mode(mode)
    o32, o16 = ops.gemm_f16(As[0], w, bias=b, residual=res, out_f32=True, out_f16=True)
    torch.cuda.synchronize()
    print(f"mode {mode}: max |out32 - ref| = {(o32 - ref).abs().max().item():.3e}, |out16 - ref| = {(o16.float() - ref).abs().max().item():.3e}")
for mode in (128, 256):
    run("f16 out only", mode, out_f32=False, out_f16=True)
    run("fp32 residual in, fp32 out", mode, residual=res, out_f32=True)
    run("fp32 residual in, fp32 + f16 out", mode, residual=res, out_f32=True, out_f16=True)
