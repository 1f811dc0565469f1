"""Thin tensor-level wrappers over the C ABI (device tensors in, device tensors out).
Used by the nn.Module mirrors in ``modules.py`` and by the parity tests."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from ._lib import c_float, c_int, ptr, stream


def _chk_f16(t, name):
    assert t.is_cuda and t.dtype == torch.float16 and t.stride(-1) == 1, f"{name}: need contiguous-cuda-f16"


def gemm_f16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
             residual: Optional[torch.Tensor] = None, out_f32: bool = True, out_f16: bool = False):
    """C = a @ w.T (+bias)(relu)(+residual).  a [M,K] f16, w [N,K] f16.  Returns (out32|None, out16|None)."""
    L.init(a.device)
    _chk_f16(a, "a"); _chk_f16(w, "w")
    M, K = a.shape
    N = w.shape[0]
    o32 = torch.empty(M, N, device=a.device, dtype=torch.float32) if out_f32 else None
    o16 = torch.empty(M, N, device=a.device, dtype=torch.float16) if out_f16 else None
    L.call("tocvp_gemm_f16", ptr(a), c_int(a.stride(0)), ptr(w), c_int(w.stride(0)), c_int(M), c_int(N), c_int(K),
           ptr(bias), c_int(int(relu)), ptr(residual), c_int(residual.stride(0) if residual is not None else 0),
           ptr(o32), c_int(N), ptr(o16), c_int(N), L.tuning_ptr(), stream())
    return o32, o16


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              add: Optional[torch.Tensor] = None, out_f16: bool = True, out_f32: bool = False):
    """Row LayerNorm of a 2-D tensor [rows, D] (fp32 or f16 in)."""
    L.init(x.device)
    rows, D = x.shape
    o16 = torch.empty(rows, D, device=x.device, dtype=torch.float16) if out_f16 else None
    o32 = torch.empty(rows, D, device=x.device, dtype=torch.float32) if out_f32 else None
    L.call("tocvp_layernorm", ptr(x), c_int(int(x.dtype == torch.float16)), c_int(x.stride(0)), ptr(add),
           c_int(add.shape[0] if add is not None else 0), ptr(gamma), ptr(beta), c_float(eps), c_int(rows), c_int(D),
           ptr(o16), c_int(D), ptr(o32), c_int(D), stream())
    return o16, o32


def cast_f16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> f16 (saturating) with the library's own kernel."""
    L.init(x.device)
    x = x.float().contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    L.call("tocvp_cast_f16", ptr(x), ptr(out), L.c_size_t(x.numel()), stream())
    return out


def mha_f16(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, Tq: int, Tk: int, heads: int) -> torch.Tensor:
    """q [B*Tq, ldq], k / v [B*Tk, ldkv] f16 views (head h at columns h*64..h*64+63, last stride 1; k and v share their row
    stride) -> out [B*Tq, heads*64] f16.  Head dim 64, Tk <= 128."""
    L.init(q.device)
    assert q.dtype == k.dtype == v.dtype == torch.float16 and k.stride(0) == v.stride(0)
    out = torch.empty(B * Tq, heads * 64, device=q.device, dtype=torch.float16)
    L.call("tocvp_mha_f16", ptr(q), c_int(q.stride(0)), ptr(k), ptr(v), c_int(k.stride(0)), c_int(B), c_int(Tq), c_int(Tk),
           c_int(heads), ptr(out), c_int(out.stride(0)), stream())
    return out


def add_table(x: torch.Tensor, table: torch.Tensor, div: int, mod: int) -> torch.Tensor:
    """out[row] = x[row] + table[(row / div) % mod] over fp32 rows of D = table.shape[-1] floats."""
    L.init(x.device)
    x = x.float().contiguous()
    D = table.shape[-1]
    out = torch.empty_like(x)
    L.call("tocvp_add_table", ptr(x), ptr(table), c_int(div), c_int(mod), c_int(D), L.c_size_t(x.numel() // D), ptr(out),
           stream())
    return out


def probe_shifted_operand(x: torch.Tensor, w: torch.Tensor, shift: int, base_offset_mode: int):
    L.init(x.device)
    out = torch.zeros(128, 64, device=x.device, dtype=torch.float32)
    rc = L.load_probe().tocvp_probe_shifted_operand(ptr(x), ptr(w), ptr(out), c_int(shift), c_int(base_offset_mode),
                                                    stream())
    if rc != 0:
        raise L.TocvpError(f"probe failed: {rc}")
    return out


def pack_conv5x5_weight(w: torch.Tensor) -> torch.Tensor:
    """torch conv weight [Cout,Cin,5,5] -> tap-major f16 [25, Cout, Cin] (tap = ky*5 + kx)."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous().half()


def conv5x5_f16(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, relu: bool = True):
    """x f16 NHWC [N,H,W,Cin] -> f16 NHWC [N,H,W,Cout]."""
    L.init(x.device)
    _chk_f16(x, "x")
    n, h, w_, ci = x.shape
    co = w_packed.shape[1]
    out = torch.empty(n, h, w_, co, device=x.device, dtype=torch.float16)
    L.call("tocvp_conv5x5_f16", ptr(x), ptr(w_packed), ptr(bias), ptr(out), c_int(n), c_int(h), c_int(w_), c_int(ci),
           c_int(co), c_int(int(relu)), L.tuning_ptr(), stream())
    return out


def set_gemm_mode(mode: int):
    """0 = automatic, 1 = single-CTA GEMM kernel only, 128 / 256 = CTA-pair kernel with that tile width; 258 / 259 =
    W-resident variant of the 256-wide pair kernel off / on (tests, tuning).  Sets the caller-owned L.TUNING."""
    if mode in (258, 259):
        L.TUNING.gemm_no_wres = int(mode == 258)
        return
    if mode not in (0, 1, 128, 256):
        raise ValueError(f"gemm mode {mode}")
    L.TUNING.gemm_mode = mode


def set_conv_mode(mode: int):
    """0 = CTA-pair conv kernel when applicable (default), 1 = single-CTA kernel only (tests, tuning)."""
    if mode not in (0, 1):
        raise ValueError(f"conv mode {mode}")
    L.TUNING.conv_mode = mode


def set_tuning(**fields):
    """Set fields of the caller-owned tocvp_tuning the modules pass with every call (encode_mode, decode_mode,
    corrector_mode, no_pdl, no_tile_alternation, gemm_mode, gemm_no_wres, conv_mode)."""
    for k, v in fields.items():
        if not hasattr(L.TUNING, k):
            raise AttributeError(k)
        setattr(L.TUNING, k, int(v))
