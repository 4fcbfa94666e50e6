"""Packed-weight layouts (hoig_b200/packing.py) against plain torch convolutions, through the CPU op emulation."""
import torch
import torch.nn.functional as F

from hoig_b200.packing import pack_conv_weight, pack_spade_gamma_beta, pack_unfolded3_weight

from . import emu_ops


def test_unfolded3_weight_matches_a_3x3_conv_on_the_resized_segmap():
    g = torch.Generator().manual_seed(0)
    seg = torch.randn(2, 12, 32, 32, generator=g)
    w = torch.randn(128, 12, 3, 3, generator=g) * 0.1
    b = torch.randn(128, generator=g) * 0.1
    u = emu_ops.seg_unfold3(seg, torch.empty(2, 16, 16, 128))
    out = emu_ops.conv2d(u, pack_unfolded3_weight(w, torch.float32), torch.empty(2, 16, 16, 128), kh=1, kw=1, bias=b, act=1)
    ref = F.relu(F.conv2d(F.interpolate(seg, size=(16, 16), mode="nearest"), w, b, padding=1)).permute(0, 2, 3, 1)
    assert (out - ref).abs().max().item() <= 1e-5


def test_interleaved_gamma_beta_rows_are_a_permutation_of_the_stacked_ones():
    g = torch.Generator().manual_seed(1)
    c, hid = 24, 16
    wg, wb = torch.randn(c, hid, 3, 3, generator=g), torch.randn(c, hid, 3, 3, generator=g)
    bg, bb = torch.randn(c, generator=g), torch.randn(c, generator=g)
    ws, bs = pack_spade_gamma_beta(wg, bg, wb, bb, torch.float32)
    wi, bi = pack_spade_gamma_beta(wg, bg, wb, bb, torch.float32, interleave=True)
    for ch in range(c):
        blk, j = divmod(ch, 8)
        assert torch.equal(wi[blk * 16 + j], ws[ch]) and torch.equal(wi[blk * 16 + 8 + j], ws[c + ch])
        assert bi[blk * 16 + j] == bs[ch] and bi[blk * 16 + 8 + j] == bs[c + ch]


def test_transposed_pack_reproduces_conv_transpose():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 6, 5, 16, generator=g)                       # NHWC, Cin 16
    w = torch.randn(16, 8, 3, 3, generator=g) * 0.2                 # ConvTranspose2d weight (Cin, Cout, 3, 3)
    out = emu_ops.conv2d(x, pack_conv_weight(w, torch.float32, transposed=True), torch.empty(1, 12, 10, 8), kh=3, kw=3, stride=2, pad=1, mode=1)
    ref = F.conv_transpose2d(x.permute(0, 3, 1, 2), w, stride=2, padding=1, output_padding=1).permute(0, 2, 3, 1)
    assert (out - ref).abs().max().item() <= 1e-5
