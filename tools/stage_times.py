"""CUDA-event timing of the four stages of one rollout step (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import rollout, weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev)
videos, text, noise = weights.synthetic_inputs(B, 20, 32, seed=0)
videos, text = videos.to(dev), text.to(dev)
ssd = weights.savi_state_dict(14)
init = (ssd["initializer.slots_mu"] + ssd["initializer.slots_sigma"] * noise).to(dev)


def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r

ms_dec, out = t(lambda: savi(mode="decomp", x=videos, num_imgs=20, decode=False, init_slots=init))
sh = out["slot_history"]
xs = videos.reshape(B * 20, 3, 64, 64)
ms_enc, (f16, _) = t(lambda: savi._encode_raw(xs, B * 20, 3 * 64 * 64, False))
ms_pred, ps = t(lambda: pred(sh, text_embeddings=text))
ms_decode, _ = t(lambda: savi.decode(ps.reshape(B * 19, 8, 128)))
ms_decode_img, _ = t(lambda: savi.decode(ps.reshape(B * 19, 8, 128), only_imgs=True))
cur = init.clone(); o = torch.empty_like(cur); nx = torch.empty_like(cur)
ms_sa3, _ = t(lambda: savi.slot_attention.run(f16, 20 * 4096 * 128, B, 4096, cur, 3, o, 8 * 128, nx))
ms_sa1, _ = t(lambda: savi.slot_attention.run(f16, 20 * 4096 * 128, B, 4096, cur, 1, o, 8 * 128, nx))
feats32 = torch.randn(B, 4096, 128, device=dev)
ms_sa3_f32, _ = t(lambda: savi.slot_attention.run(feats32, 4096 * 128, B, 4096, cur, 3, o, 8 * 128, None))
print(f"B={B}: decomp(20 frames) {ms_dec:.1f} ms [encode {ms_enc:.1f}, SA 3it {ms_sa3:.2f}, SA 1it+trans {ms_sa1:.2f}] | "
      f"predict {ms_pred:.1f} ms | decode {ms_decode:.1f} ms (imgs only {ms_decode_img:.1f})")
print(f"SA microbench (config 2) fp32 feats, 3 iters: {ms_sa3_f32:.3f} ms -> {B*4096*128*4/ms_sa3_f32/1e6:.0f} GB/s algorithmic "
      f"(read-once); f16 feats: {ms_sa3:.3f} ms -> {B*4096*128*2/ms_sa3/1e6:.0f} GB/s")
