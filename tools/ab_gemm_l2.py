"""Is the K = 2048 / N = 512 projection limited by WHERE its A operand comes from?  Same shape and tile count (M = 9472 =
37 x 256 rows: 148 tiles of 256 x 128, two full waves), A either L2-resident (one 38.8 MB buffer re-used) or streamed from
DRAM (rotation over 5 buffers = 194 MB).  Dev tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import ops
from tools.microbench import timeit
for (M, N, K) in [(9472, 512, 2048), (9472, 2048, 512), (9472, 512, 512)]:
    bufs = [torch.randn(M, K, device="cuda").half() for _ in range(5 if K == 2048 else 20)]
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    state = {"i": 0}
    def rot():
        state["i"] = (state["i"] + 1) % len(bufs)
        return ops.gemm_f16(bufs[state["i"]], w, out_f32=False, out_f16=True)
    for mode in (128, 256):
        ops.set_gemm_mode(mode)
        t_l2 = timeit(lambda: ops.gemm_f16(bufs[0], w, out_f32=False, out_f16=True), iters=40)
        t_dr = timeit(rot, iters=40)
        fl = 2 * M * N * K
        print(f"{M}x{N}x{K} BN={mode}: A in L2 {t_l2*1e3:6.1f} us ({fl/t_l2/1e9:6.0f} TF) | A from DRAM {t_dr*1e3:6.1f} us "
              f"({fl/t_dr/1e9:6.0f} TF)", flush=True)
ops.set_gemm_mode(0)
