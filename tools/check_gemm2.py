import sys; sys.path.insert(0, "/root/repo")
import torch
from textocvp_b200 import ops
def rel(a, b):
    a, b = a.double(), b.double(); return float((a - b).norm() / b.norm())
torch.manual_seed(0)
for mode in (128, 256):
    ops.set_gemm_mode(mode)
    for M, N, K in [(256, 256, 64), (512, 512, 512), (20480, 512, 512), (1000, 2048, 512), (4096, 512, 2048), (2048, 1536, 512), (300, 768, 128)]:
        if N % mode: continue
        a = torch.randn(M, K, device="cuda").half()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        bias = torch.randn(N, device="cuda"); res = torch.randn(M, N, device="cuda")
        ref = a.double() @ w.double().t()
        o32, o16 = ops.gemm_f16(a, w, out_f32=True, out_f16=True)
        torch.cuda.synchronize()
        e1 = rel(o32, ref)
        o32b, _ = ops.gemm_f16(a, w, bias=bias, relu=True, residual=res)
        torch.cuda.synchronize()
        e2 = rel(o32b, torch.relu(ref + bias.double()) + res.double())
        print(f"mode {mode} {M}x{N}x{K}: err {e1:.2e} {e2:.2e} f16 {rel(o16, ref):.2e}", flush=True)
ops.set_gemm_mode(0)
