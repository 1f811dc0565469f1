import os, sys
sys.path.insert(0, "/root/repo")
import torch
from textocvp_b200 import ops
N, K = 512, 2048
w = (torch.randn(N, K, device="cuda") / 45).half()
b = torch.randn(N, device="cuda")
for mode in (128, 256):
    ops.set_gemm_mode(mode)
    for rounds in (1, 2, 3, 4):
        M = 256 * 37 * rounds
        As = [torch.randn(M, K, device="cuda").half() for _ in range(3)]
        for i in range(3): ops.gemm_f16(As[i % 3], w, bias=b, out_f32=False, out_f16=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for i in range(n): ops.gemm_f16(As[i % 3], w, bias=b, out_f32=False, out_f16=True)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        print(f"mode {mode} M {M:6d} ({rounds} x 74 tiles of 256x256): {us:7.1f} us  {2.0*M*N*K/us/1e6:7.1f} TF")
        del As
