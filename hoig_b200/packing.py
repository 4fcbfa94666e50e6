"""Weight packing for the implicit-GEMM kernels (host-side plumbing, plain torch ops).

Checkpoints keep the reference's OIHW fp32 ``nn.Parameter`` layout; the kernels
consume a K-major matrix ``Wp[Cout_pad16][cols]``:

* ``nn.Conv2d`` (and the local-attention k5s5 conv): ``k = (r*KW + s)*Cin_pad + c``, padded to 64 columns
  (``Cin_pad`` = Cin rounded up to 8, matching the zero-padded NHWC activations);
* ``nn.ConvTranspose2d`` (stride 2): the four output-parity phases (a,b) side by side, each
  ``[taps(a) x taps(b) x Cin_pad]`` padded to 64 columns, where kernel row ``r`` belongs to parity ``a`` iff
  ``(a + pad - r)`` is even (it then reads input row ``gy + (a + pad - r)/2``) -- see ``csrc/conv_plan.cu``.

Packed copies are derived caches.
"""
from __future__ import annotations

import torch


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def transposed_axis_taps(k: int, pad: int):
    """Per output parity a in (0,1): kernel indices r with (a + pad - r) even, in increasing order."""
    return [[r for r in range(k) if (a + pad - r) % 2 == 0] for a in (0, 1)]


def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype, transposed: bool = False, pad: int = 1) -> torch.Tensor:
    """``nn.Conv2d.weight`` (Cout,Cin,KH,KW) or ``nn.ConvTranspose2d.weight`` (Cin,Cout,KH,KW) -> packed matrix."""
    w = w.detach().float()
    if transposed:
        w = w.permute(1, 0, 2, 3)
    cout, cin, kh, kw = w.shape
    cin_p = ceil_to(cin, 8)
    rows = ceil_to(cout, 16)
    t = torch.zeros(cout, kh, kw, cin_p, dtype=torch.float32, device=w.device)
    t[..., :cin] = w.permute(0, 2, 3, 1)
    if not transposed:
        groups = [[(r, s) for r in range(kh) for s in range(kw)]]
    else:
        ra, sa = transposed_axis_taps(kh, pad), transposed_axis_taps(kw, pad)
        groups = [[(r, s) for r in ra[a] for s in sa[b]] for a in (0, 1) for b in (0, 1)]
    blocks = []
    for taps in groups:
        n = len(taps) * cin_p
        blk = torch.zeros(rows, ceil_to(n, 64), dtype=torch.float32, device=w.device)
        if taps:
            blk[:cout, :n] = torch.stack([t[:, r, s] for r, s in taps], 1).reshape(cout, n)
        blocks.append(blk)
    return torch.cat(blocks, 1).to(dtype).contiguous()


def pack_spade_gamma_beta(wg, bg, wb, bb, dtype):
    """mlp_gamma / mlp_beta (spade.py:22-23) fused into one GEMM with N = 2C:
    rows [0,C) produce gamma, rows [C,2C) beta."""
    w = torch.cat([wg, wb], 0)
    b = torch.cat([bg, bb], 0).detach().float().contiguous()
    return pack_conv_weight(w, dtype), b
