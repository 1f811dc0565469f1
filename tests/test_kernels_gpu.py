"""GPU parity tests of the building-block kernels through the C ABI (run with -m gpu on a B200).
Floating point -> compared against a plain torch fp32/fp64 reference of the same op; tolerances are
written next to each check."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 512, 512), (2048, 1536, 512), (1000, 2048, 512),
                                   (4096, 512, 2048), (520, 128, 32), (333 * 8, 1600, 128), (64, 64, 128),
                                   (20480, 512, 512)])
def test_gemm_f16(M, N, K):
    from textocvp_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    ref = a.double() @ w.double().t()
    o32, o16 = ops.gemm_f16(a, w, out_f32=True, out_f16=True)
    torch.cuda.synchronize()
    assert _rel(o32, ref) < 1e-5          # exact f16 products, fp32 accumulation
    assert _rel(o16, ref) < 1e-3          # + f16 output rounding
    o32, _ = ops.gemm_f16(a, w, bias=bias, relu=True, residual=res)
    ref2 = torch.relu(ref + bias.double()) + res.double()
    assert _rel(o32, ref2) < 1e-5


@pytest.mark.parametrize("mode", [128, 256])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (20480, 512, 512), (1000, 2048, 512), (4096, 512, 2048),
                                   (300, 768, 128), (2048, 1536, 512), (1300, 776, 1024), (700, 264, 64)])
def test_gemm_pair_kernel(mode, M, N, K):
    """CTA-pair (cta_group::2) kernel forced on, both tile widths: ragged M tails, both outputs at once (TMA-store
    epilogue through shared staging tiles), bias + ReLU + residual."""
    from textocvp_b200 import ops
    if N < mode:
        pytest.skip("narrower than one tile")
    g = torch.Generator(device="cuda").manual_seed(M + N + K + mode)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    ref = a.double() @ w.double().t()
    ops.set_gemm_mode(mode)
    try:
        o32, o16 = ops.gemm_f16(a, w, out_f32=True, out_f16=True)
        o32b, _ = ops.gemm_f16(a, w, bias=bias, relu=True, residual=res)
        _, o16c = ops.gemm_f16(a, w, bias=bias, relu=True, out_f32=False, out_f16=True)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_mode(0)
    assert _rel(o32, ref) < 1e-5
    assert _rel(o16, ref) < 1e-3
    assert _rel(o32b, torch.relu(ref + bias.double()) + res.double()) < 1e-5
    assert _rel(o16c, torch.relu(ref + bias.double())) < 1e-3


@pytest.mark.parametrize("rows,D,f16", [(1000, 512, False), (77, 128, False), (4096, 32, True), (300, 768, False)])
def test_layernorm(rows, D, f16):
    from textocvp_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 1
    if f16:
        x = x.half()
    gamma = torch.randn(D, device="cuda", generator=g)
    beta = torch.randn(D, device="cuda", generator=g)
    add = torch.randn(64, D, device="cuda", generator=g)
    o16, o32 = ops.layernorm(x, gamma, beta, 1e-6, out_f32=True)
    ref = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-6)
    assert _rel(o32, ref) < 1e-5
    assert _rel(o16, ref) < 1e-3
    _, o32 = ops.layernorm(x, gamma, beta, 1e-3, add=add, out_f16=False, out_f32=True)
    xa = x.double() + add.double().repeat((rows + 63) // 64, 1)[:rows]
    ref = torch.nn.functional.layer_norm(xa, (D,), gamma.double(), beta.double(), 1e-3)
    assert _rel(o32, ref) < 1e-5


@pytest.mark.parametrize("shift", [0, 1, 3, 8, 13, 68, 70, 127])
def test_probe_shifted_sw128_operand(shift):
    """The conv kernel's assumption: a SW128 K-major operand may start at any 128-byte row."""
    from textocvp_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(shift)
    x = torch.randn(256, 64, device="cuda", generator=g).half()
    w = torch.randn(64, 64, device="cuda", generator=g).half()
    out = ops.probe_shifted_operand(x, w, shift, 0)
    ref = x[shift:shift + 128].double() @ w.double().t()
    assert _rel(out, ref) < 1e-5


@pytest.mark.parametrize("n,c", [(1, 64), (3, 64), (40, 64), (2, 32), (37, 32)])
@pytest.mark.parametrize("mode", [0, 1])
def test_conv5x5(n, c, mode):
    """tcgen05 implicit-GEMM conv (mode 0: CTA-pair kernel when the tile count is even, mode 1: single-CTA kernel)
    vs torch conv2d (fp32) on f16-rounded operands."""
    from textocvp_b200 import ops
    ops.set_conv_mode(mode)
    g = torch.Generator(device="cuda").manual_seed(n * 100 + c)
    x = torch.randn(n, 64, 64, c, device="cuda", generator=g).half()
    w = (torch.randn(c, c, 5, 5, device="cuda", generator=g) / (25 * c) ** 0.5)
    b = torch.randn(c, device="cuda", generator=g)
    try:
        out = ops.conv5x5_f16(x, ops.pack_conv5x5_weight(w), b, relu=True)
        torch.cuda.synchronize()
    finally:
        ops.set_conv_mode(0)
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), b, padding=2))
    ref = ref.permute(0, 2, 3, 1)
    assert _rel(out, ref) < 1e-3      # f16 output rounding only (operands identical, fp32 accumulate)
    assert (out.float() - ref).abs().max() < 2e-2


@pytest.mark.parametrize("B,Tq,Tk,cross", [(3, 80, 80, False), (2, 8, 8, False), (5, 24, 24, False), (4, 80, 32, True),
                                           (2, 72, 16, True), (2, 100, 100, False)])
def test_mha(B, Tq, Tk, cross):
    """mma.sync attention vs torch (fp32 softmax) on the same f16 q/k/v."""
    from textocvp_b200 import _lib as L
    from textocvp_b200._lib import c_int, ptr, stream
    H, dh = 8, 64
    g = torch.Generator(device="cuda").manual_seed(Tq * 7 + Tk)
    if cross:
        q = torch.randn(B * Tq, H * dh, device="cuda", generator=g).half()
        kv = torch.randn(B * Tk, 2 * H * dh, device="cuda", generator=g).half()
        k, v, ldq, ldkv = kv, kv[:, H * dh:], H * dh, 2 * H * dh
    else:
        qkv = torch.randn(B * Tq, 3 * H * dh, device="cuda", generator=g).half()
        q, k, v, ldq, ldkv = qkv, qkv[:, H * dh:], qkv[:, 2 * H * dh:], 3 * H * dh, 3 * H * dh
    out = torch.empty(B * Tq, H * dh, device="cuda", dtype=torch.float16)
    L.init(out.device)
    L.call("tocvp_mha_f16", ptr(q), c_int(ldq), ptr(k), ptr(v), c_int(ldkv), c_int(B), c_int(Tq), c_int(Tk), c_int(H),
           ptr(out), c_int(H * dh), stream())
    qf = q[:, :H * dh].double().view(B, Tq, H, dh).transpose(1, 2)
    kf = k[:, :H * dh].double().view(B, Tk, H, dh).transpose(1, 2)
    vf = v[:, :H * dh].double().view(B, Tk, H, dh).transpose(1, 2)
    ref = ((qf @ kf.transpose(-1, -2)) * dh ** -0.5).softmax(-1) @ vf
    ref = ref.transpose(1, 2).reshape(B * Tq, H * dh)
    assert _rel(out, ref) < 2e-3      # P and the output are rounded to f16


@pytest.mark.parametrize("B,F,H,W,T,f0", [(3, 4, 64, 64, 6, 1), (2, 3, 128, 128, 4, 1), (1, 2, 40, 72, 2, 0)])
def test_frame_metrics(B, F, H, W, T, f0):
    """clamp + MSE / PSNR / SSIM kernels vs the oracle's restatement of piqa 1.2.2 (tolerance: fp32 reductions, 1e-4
    relative on MSE/SSIM, 1e-3 dB on PSNR); targets are read in place from the video tensor."""
    from oracle import textocvp_oracle as O
    from textocvp_b200 import rollout
    g = torch.Generator().manual_seed(B * 100 + H)
    videos = torch.rand(B, T, 3, H, W, generator=g)
    pred = videos[:, f0:f0 + F] + 0.3 * torch.randn(B, F, 3, H, W, generator=g)       # leaves [0,1]: clamp matters
    m = rollout.frame_metrics(pred.cuda(), videos.cuda(), f0)
    p, t = pred.clamp(0, 1), videos[:, f0:f0 + F].clamp(0, 1)
    mse = ((p - t) ** 2).flatten(2).mean(-1)
    assert _rel(m["mse"].cpu(), mse) < 1e-4
    assert (m["psnr"].cpu() - O.psnr(p, t)).abs().max() < 1e-3
    assert (m["ssim"].cpu() - O.ssim(p, t)).abs().max() < 1e-4


def test_lpips_hook_feeds_metric_sums():
    """LPIPS is a hook (its AlexNet weights cannot be fetched offline): any callable on the clamped frames flows into the
    accumulators the way piqa's LPIPS does in the reference (lib/metrics.py:266-298)."""
    from textocvp_b200 import rollout
    g = torch.Generator().manual_seed(9)
    pred = (torch.rand(3, 4, 3, 64, 64, generator=g) * 1.4 - 0.2).cuda()
    vid = torch.rand(3, 6, 3, 64, 64, generator=g).cuda()
    fake = lambda a, b: (a - b).abs().flatten(1).mean(1)
    m = rollout.frame_metrics(pred, vid, 1, lpips_fn=fake)
    ref = (pred.clamp(0, 1) - vid[:, 1:5].clamp(0, 1)).abs().flatten(2).mean(2)
    assert m["lpips"].shape == (3, 4) and torch.allclose(m["lpips"], ref, atol=1e-6)
    sums = rollout.MetricSums(4, pred.device)
    sums.accumulate(m["psnr"], m["mse"], m["ssim"], m["lpips"])
    assert abs(sums.results()["lpips_mean"] - float(ref.mean())) < 1e-5
