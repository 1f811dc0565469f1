"""Host-side weight-packing algebra (textocvp_b200/modules.py), checked on the CPU in fp64-ish torch against the plain
formulation the reference executes: BatchNorm folding, Upsample(2)->conv3x3 as four 2x2 phase convolutions on the
low-resolution input (the flattened zero-bordered layout the CUDA implicit GEMM walks), the narrow final conv, and
LayerNorm folded into the consuming projection."""
import torch
import torch.nn.functional as F

from textocvp_b200 import modules as M


def _conv_from_packed_phase(x, packed, bias, co):
    """Emulate the CUDA kernel's indexing: zero-bordered low-res input, per-phase 2x2 taps with offsets
    (dy+py-1, dx+px-1), output pixel (2y+py, 2x+px)."""
    n, ci, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))
    out = x.new_zeros(n, co, 2 * H, 2 * W)
    wp = packed.reshape(4, co, 4, ci)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        acc = x.new_zeros(n, co, H, W)
        for t in range(4):
            dy, dx = t >> 1, t & 1
            oy, ox = dy + py - 1, dx + px - 1
            win = xp[:, :, 1 + oy:1 + oy + H, 1 + ox:1 + ox + W]
            acc += torch.einsum("nchw,oc->nohw", win, wp[ph, :, t])
        out[:, :, py::2, px::2] = acc + bias[ph * co:(ph + 1) * co].view(1, -1, 1, 1)
    return out


def test_phase_convolution_equals_upsample_then_conv():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 5, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 6, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(4, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)
    out = _conv_from_packed_phase(x, M.pack_conv3x3_phase(w), b.repeat(4), 4)
    assert torch.allclose(out, ref, atol=1e-12)


def test_final_conv_phase_packing():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 8, 6, 6, generator=g, dtype=torch.float64)
    w = torch.randn(3, 8, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)
    packed, bias = M.pack_final_conv_phase(w, b)
    assert packed.shape == (64, 72) and (packed[16:] == 0).all()
    wp = packed[:16].reshape(4, 4, 9, 8)                      # [phase, c(3+pad), tap, ci]
    xp = F.pad(x, (1, 1, 1, 1))
    out = x.new_zeros(1, 3, 12, 12)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        acc = x.new_zeros(1, 4, 6, 6)
        for tap in range(9):
            oy, ox = tap // 3 - 1, tap % 3 - 1
            acc += torch.einsum("nchw,oc->nohw", xp[:, :, 1 + oy:7 + oy, 1 + ox:7 + ox], wp[ph, :, tap])
        out[:, :, py::2, px::2] = (acc + bias[ph * 4:ph * 4 + 4].view(1, 4, 1, 1))[:, :3]
    assert torch.allclose(out, ref, atol=1e-12)
    assert (wp[:, 3] == 0).all() and (bias[3::4] == 0).all()  # the pad channel stays zero


def test_plain_packing_and_batchnorm_fold():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 5, 4, 4, generator=g)
    w, b = torch.randn(7, 5, 3, 3, generator=g), torch.randn(7, generator=g)
    bw, bb = torch.rand(7, generator=g) + 0.5, torch.randn(7, generator=g)
    mean, var = torch.randn(7, generator=g), torch.rand(7, generator=g) + 0.5
    ref = F.batch_norm(F.conv2d(x, w, b, padding=1), mean, var, bw, bb, False, 0.0, 1e-5)
    wf, bf = M.fold_batchnorm(w, b, bw, bb, mean, var, 1e-5)
    assert torch.allclose(F.conv2d(x, wf, bf, padding=1), ref, atol=1e-5)
    packed = M.pack_conv3x3_plain(wf)                           # K = (ky*3+kx)*ci + c
    cols = F.unfold(x, 3, padding=1).reshape(2, 5, 9, 16).permute(0, 3, 2, 1).reshape(2, 16, 45)   # [n, pix, tap*ci]
    out = (cols @ packed.t() + bf).permute(0, 2, 1).reshape(2, 7, 4, 4)
    assert torch.allclose(out, ref, atol=1e-4)


def test_layernorm_fold():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(50, 64, generator=g) * 2 + 0.3
    w, bias = torch.randn(24, 64, generator=g) / 8, torch.randn(24, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(64, generator=g), 0.1 * torch.randn(64, generator=g)
    ref = F.linear(F.layer_norm(x, (64,), gamma, beta, 1e-6), w, bias)
    wf, c, d = M.fold_layernorm(w, gamma, beta, bias)
    mu, var = x.mean(-1, keepdim=True), x.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-6)
    out = rstd * (x @ wf.float().t() - mu * c) + d              # what the consumer GEMM's epilogue evaluates
    assert float((out - ref).norm() / ref.norm()) < 1e-3        # only the f16 rounding of w*gamma separates them
    exact = rstd * (x @ (w * gamma).t() - mu * (w * gamma).sum(1)) + d
    assert torch.allclose(exact, ref, atol=1e-4)


def test_conv1_vertical_pair_packing():
    """Encoder conv 1 with the x-taps folded into K: sum over 3 vertical taps of Wp[j] . P[y + 2j - 2] on the x-im2col input
    (modules.im2col_x_row_pairs = what enc_pack_vp_kernel writes) equals the 5x5 convolution with zero padding."""
    g = torch.Generator().manual_seed(3)
    w = torch.randn(32, 3, 5, 5, generator=g)
    b = torch.randn(32, generator=g)
    x = torch.rand(2, 3, 16, 32, generator=g)
    ref = F.conv2d(x, w, b, padding=2)                                          # [n, 32, H, W]
    wp = M.pack_conv1_vertical_pairs(w)                                          # [3, 32, 32]
    P = M.im2col_x_row_pairs(x)                                                  # [n, H+1, W, 32], stored row r = image row r-1
    assert wp.shape == (3, 32, 32) and P.shape == (2, 17, 32, 32)
    assert (wp[:, :, 15] == 0).all() and (wp[:, :, 31] == 0).all() and (wp[2, :, 16:] == 0).all()
    H = x.shape[2]
    Pz = F.pad(P, (0, 0, 0, 0, 2, 2))                                            # stored rows -2 .. H+2 (TMA zero fill)
    out = b.view(1, 1, 1, -1).expand(2, H, 32, 32).clone()
    for j in range(3):
        rows = Pz[:, 2 * j + 1:2 * j + 1 + H]                                    # stored row (y + 2j - 2) + 1, offset by the pad 2
        out = out + torch.einsum("nyxk,ok->nyxo", rows, wp[j])
    assert torch.allclose(out.permute(0, 3, 1, 2), ref, atol=1e-4, rtol=1e-4)


def test_conv_x_pair_packing():
    """Encoder 32 -> 32 layers in pixel-pair form (conv5x5_tc.cu, XP): with one MMA row = pixels (2j, 2j+1) and the six input
    shifts u = 0..5 of a filter row, block (ky, u) of modules.pack_conv_x_pairs applied to input pixel 2j + u - 2 and summed
    over (ky, u) gives both pixels of the pair of the zero-padded 5x5 convolution."""
    g = torch.Generator().manual_seed(5)
    w = torch.randn(32, 32, 5, 5, generator=g)
    x = torch.randn(2, 32, 16, 64, generator=g)
    ref = F.conv2d(x, w, padding=2)
    wp = M.pack_conv_x_pairs(w)
    assert wp.shape == (30, 64, 32)
    assert (wp[0::6, 32:] == 0).all() and (wp[5::6, :32] == 0).all()           # the taps that do not exist are zero blocks
    n, c, H, W = x.shape
    xz = F.pad(x, (2, 4, 2, 2))                                                 # columns -2 .. W+3 (TMA zero fill)
    out = torch.zeros(n, 32, H, W)
    for ky in range(5):
        for u in range(6):
            inp = xz[:, :, ky:ky + H, u:u + W:2]                                # input pixel 2j + u - 2 of row y + ky - 2
            res = torch.einsum("oc,nchw->nohw", wp[ky * 6 + u], inp)            # [n, 64, H, W/2]
            out[:, :, :, 0::2] += res[:, :32]
            out[:, :, :, 1::2] += res[:, 32:]
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4)
