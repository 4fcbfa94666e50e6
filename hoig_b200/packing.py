"""Weight packing for the implicit-GEMM kernels (host-side plumbing, plain torch ops).

Checkpoints keep the reference's OIHW fp32 ``nn.Parameter`` layout; the kernels
consume a K-major matrix ``Wp[Cout_pad16][K_pad64]`` with
``k = (r*KW + s)*Cin_pad + c`` (``Cin_pad`` = Cin rounded up to 8, matching the
zero-padded NHWC activations).  Packed copies are derived caches.
"""
from __future__ import annotations

import torch


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype, transposed: bool = False) -> torch.Tensor:
    """``nn.Conv2d.weight`` (Cout,Cin,KH,KW) or ``nn.ConvTranspose2d.weight`` (Cin,Cout,KH,KW)
    -> packed (Cout_pad16, K_pad64).  For the transposed case tap (r,s) multiplies
    input pixel ((oy+pad-r)/stride, (ox+pad-s)/stride), i.e. no kernel flip is needed."""
    if transposed:
        w = w.permute(1, 0, 2, 3)
    cout, cin, kh, kw = w.shape
    cin_p = ceil_to(cin, 8)
    rows, cols = ceil_to(cout, 16), ceil_to(kh * kw * cin_p, 64)
    out = torch.zeros(rows, cols, dtype=torch.float32, device=w.device)
    t = torch.zeros(cout, kh, kw, cin_p, dtype=torch.float32, device=w.device)
    t[..., :cin] = w.detach().float().permute(0, 2, 3, 1)
    out[:cout, : kh * kw * cin_p] = t.reshape(cout, -1)
    return out.to(dtype).contiguous()


def pack_spade_gamma_beta(wg, bg, wb, bb, dtype):
    """mlp_gamma / mlp_beta (spade.py:22-23) fused into one GEMM with N = 2C:
    rows [0,C) produce gamma, rows [C,2C) beta."""
    w = torch.cat([wg, wb], 0)
    b = torch.cat([bg, bb], 0).detach().float().contiguous()
    return pack_conv_weight(w, dtype), b
