"""Three predictor GEMM shapes, a few launches each (ncu target). Dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import ops
M = 20480
a512 = torch.randn(M, 512, device="cuda").half()
a2048 = torch.randn(M, 2048, device="cuda").half()
w1 = (torch.randn(2048, 512, device="cuda") / 23).half()
w2 = (torch.randn(512, 2048, device="cuda") / 45).half()
wo = (torch.randn(512, 512, device="cuda") / 23).half()
b1 = torch.randn(2048, device="cuda"); b2 = torch.randn(512, device="cuda")
res = torch.randn(M, 512, device="cuda")
for _ in range(3):
    ops.gemm_f16(a512, w1, bias=b1, relu=True, out_f32=False, out_f16=True)       # MLP1
    ops.gemm_f16(a2048, w2, bias=b2, residual=res, out_f32=True)                  # MLP2
    ops.gemm_f16(a512, wo, residual=res, out_f32=True)                            # out-proj
torch.cuda.synchronize()
print("done")
