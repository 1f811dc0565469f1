"""Tiny end-to-end pass of both model families and of every round-2 kernel (compute-sanitizer target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from textocvp_b200 import modules as M, ops, rollout, weights
dev = torch.device("cuda:0")
savi, pred, _ = rollout.build_models(dev, num_context=1, num_preds=2)
pred.predictor.use_cuda_graph = False
videos, text, noise = weights.synthetic_inputs(3, 3, 32, seed=1)          # odd batch: padded rows in the corrector update
out = rollout.forward_eval(savi, pred, videos.to(dev), text.to(dev), 1, 2)
dino, dpred, _ = rollout.build_dino_models(dev, num_preds=2)
dpred.predictor.use_cuda_graph = False
feats, dtext, _ = weights.synthetic_dino_inputs(2, 3, 81, L=16, seed=0)
o2 = rollout.forward_eval_dino(dino, dpred, feats.to(dev), dtext.to(dev), 1, 2)
# multi-chunk decode: the chunk-pipelined driver (side stream, three activation buffers) with a ragged tail; fused
# compositing with and without the per-slot outputs, and the separate compositing kernel
d2 = savi.decode(torch.randn(256 + 8, 8, 128, device=dev), only_imgs=True)
d3 = savi.decode(torch.randn(5, 8, 128, device=dev))
ops.set_tuning(decode_mode=16)
d4 = savi.decode(torch.randn(5, 8, 128, device=dev))
ops.set_tuning(decode_mode=0)
# stand-alone sub-module forwards
t = savi.transition_module(torch.randn(3, 8, 128, device=dev))
e = savi.encoder(torch.rand(2, 3, 64, 64, device=dev))
c = savi.decoder(torch.randn(1, 128, 64, 64, device=dev))
pe = pred.predictor.pe(torch.randn(2, 3, 8, 512, device=dev), 2, 8)
# ViT front-end: short and streaming attention
enc = M.get_vit_encoder({"encoder_name": "vit_base_patch14_dinov2", "encoder_params": {"num_blocks": 2}}, 128).to(dev).eval()
v1 = enc(torch.rand(2, 3, 128, 128, device=dev))
enc2 = M.get_vit_encoder({"encoder_name": "vit_base_patch14_dinov2", "encoder_params": {"num_blocks": 1}}, 224).to(dev).eval()
v2 = enc2(torch.rand(1, 3, 224, 224, device=dev))                        # 257 tokens -> streaming attention kernel
torch.cuda.synchronize()
print("ok", float(out["psnr"].mean()), tuple(o2["pred_imgs"].shape), tuple(v1.shape), tuple(v2.shape), tuple(c.shape))
