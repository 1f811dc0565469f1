"""Tile-width choice of the pair GEMM over the row counts the rollout meets (window 1..10 frames x 8 slots x 256 sequences):
time of the 128-wide, the 256-wide and the automatic choice for the LayerNorm-consumer shapes.  Dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from microbench import timeit
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for N, K in ((1536, 512), (2048, 512), (512, 512)):
    w = (torch.randn(N, K, device="cuda") / 23).half()
    for win in range(1, 11):
        M = B * 8 * win
        a = torch.randn(M, K, device="cuda").half()
        res = []
        for mode in (128, 256, 0):
            ops.set_gemm_mode(mode)
            us = timeit(lambda: ops.gemm_f16(a, w, out_f32=False, out_f16=True), iters=30) * 1e3
            res.append(us)
        ops.set_gemm_mode(0)
        best = min(res[0], res[1])
        flag = "" if res[2] <= best * 1.03 else "   <-- automatic choice is %.0f %% slower than the best" % (100 * (res[2] / best - 1))
        print(f"N {N:5d} K {K:4d} M {M:6d}: 128-wide {res[0]:6.1f} us | 256-wide {res[1]:6.1f} us | auto {res[2]:6.1f} us{flag}", flush=True)
