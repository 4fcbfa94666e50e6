"""Weight packing for the implicit-GEMM kernels (host-side plumbing, plain torch ops).

Checkpoints keep the reference's OIHW fp32 ``nn.Parameter`` layout; the kernels
consume a K-major matrix ``Wp[Cout_pad16][cols]``:

* ``nn.Conv2d`` (and the local-attention k5s5 conv): ``k = (r*KW + s)*Cin_pad + c``, padded to 64 columns
  (``Cin_pad`` = Cin rounded up to 8, matching the zero-padded NHWC activations);
* ``nn.ConvTranspose2d`` (k3 s2 p1 op1): ``[4*Cout][4*Cin_pad]`` -- row block = output parity (a,b) in the order
  (0,0), (0,1), (1,1), (1,0) (``parity_block``), column block (dy,dx) = input tap of the 2x2 neighbourhood; blocks a parity does
  not use are zero (``csrc/conv_plan.cu``) and are skipped by the tensor-core kernel.

Packed copies are derived caches.
"""
from __future__ import annotations

import torch


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def parity_block(a: int, b: int) -> int:
    """Row block of output parity (a, b) in the packed transposed-conv matrix: order (0,0), (0,1), (1,1), (1,0).  Tap (dy, dx) of
    the 2x2 input neighbourhood feeds parity (a, b) only if dy <= a and dx <= b, so in this order the blocks every tap feeds are
    CONTIGUOUS ([0,4), [1,3), [2,4), [2,3)) and the kernel can skip the dead (n-tile, tap) combinations (csrc/conv_umma.cu)."""
    return a * 2 + (b ^ a)


def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype, transposed: bool = False, pad: int = 1) -> torch.Tensor:
    """``nn.Conv2d.weight`` (Cout,Cin,KH,KW) or ``nn.ConvTranspose2d.weight`` (Cin,Cout,3,3; stride 2) -> packed matrix."""
    w = w.detach().float()
    if not transposed:
        cout, cin, kh, kw = w.shape
        cin_p = ceil_to(cin, 8)
        t = torch.zeros(cout, kh, kw, cin_p, dtype=torch.float32, device=w.device)
        t[..., :cin] = w.permute(0, 2, 3, 1)
        out = torch.zeros(ceil_to(cout, 16), ceil_to(kh * kw * cin_p, 64), dtype=torch.float32, device=w.device)
        out[:cout, : kh * kw * cin_p] = t.reshape(cout, -1)
        return out.to(dtype).contiguous()
    # transposed conv as ONE GEMM over the 2x2 input taps (dy,dx): row block (a,b) = output parity, and
    # out[2gy+a, 2gx+b] += in[gy+dy, gx+dx] * W[:, :, a+pad-2dy, b+pad-2dx]  whenever that kernel index exists
    cin, cout, kh, kw = w.shape
    cin_p = ceil_to(cin, 8)
    out = torch.zeros(ceil_to(4 * cout, 16), ceil_to(4 * cin_p, 64), dtype=torch.float32, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    r, s_ = a + pad - 2 * dy, b + pad - 2 * dx
                    if 0 <= r < kh and 0 <= s_ < kw:
                        ph, t = parity_block(a, b), dy * 2 + dx
                        out[ph * cout:(ph + 1) * cout, t * cin_p:t * cin_p + cin] = w[:, :, r, s_].t()
    return out.to(dtype).contiguous()


def pack_unfolded3_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """3x3 conv weight (Cout,C,3,3) for a 1x1 GEMM over ``ops.seg_unfold3`` output: column t*C + c, K padded to 64."""
    w = w.detach().float()
    cout, c, kh, kw = w.shape
    assert kh == 3 and kw == 3
    out = torch.zeros(ceil_to(cout, 16), ceil_to(9 * c, 64), dtype=torch.float32, device=w.device)
    out[:cout, : 9 * c] = w.permute(0, 2, 3, 1).reshape(cout, 9 * c)
    return out.to(dtype).contiguous()


def pack_spade_gamma_beta(wg, bg, wb, bb, dtype, interleave: bool = False):
    """mlp_gamma / mlp_beta (spade.py:22-23) fused into one GEMM with N = 2C.

    ``interleave=False``: rows [0,C) produce gamma, rows [C,2C) beta (separate ``instnorm_apply``).
    ``interleave=True``: blocks of 8 channels [g0..g7 b0..b7], the layout of the SPADE-modulating conv epilogue
    (``hoigConvDesc::spade_x``), where one thread holds gamma and beta of the same channels."""
    w = torch.cat([wg, wb], 0)
    b = torch.cat([bg, bb], 0).detach().float()
    if interleave:
        c = wg.shape[0]
        assert c % 8 == 0
        idx = torch.arange(c, device=w.device).view(c // 8, 1, 8)
        perm = torch.cat([idx, idx + c], 1).reshape(-1)          # row order g-block, b-block, g-block, ...
        w, b = w[perm], b[perm]
    return pack_conv_weight(w, dtype), b.contiguous()
