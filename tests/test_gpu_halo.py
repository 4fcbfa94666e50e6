"""Tensor-core local attention (commuted formulation): replicate padding, the halo-reuse conv over padded
rasters and the combine kernel, each against its op contract (tests/emu_ops.py), and the whole block against
the BlockExtractor formulation of extract_attn.py:19-28 restated by the oracle.

Own file: a protocol bug in the tcgen05 kernel traps instead of hanging, and must not take other tests' context down.
"""
import pytest
import torch
import torch.nn.functional as F

from hoig_b200 import _lib, ops
from hoig_b200.packing import pack_conv_weight
from oracle import generator_ref as gr

from . import emu_ops

pytestmark = pytest.mark.gpu
TOL = {torch.bfloat16: 1.2e-2, torch.float16: 2e-3}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_replicate_pad(dtype):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 9, 9, 24, generator=g).to(dtype)
    for pad in (0, 2, 4):
        out = ops.replicate_pad(x.cuda(), torch.empty(3, 9 + 2 * pad, 9 + 2 * pad, 24, dtype=dtype, device="cuda"), pad)
        ref = emu_ops.replicate_pad(x, torch.empty(3, 9 + 2 * pad, 9 + 2 * pad, 24, dtype=dtype), pad)
        assert torch.equal(out.cpu(), ref)


def _halo_case(n, hp, wp, c, k, cout, dtype, variant, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, hp, wp, c, generator=g).to(dtype)
    w = torch.randn(cout, c, k, k, generator=g) * (1.0 / (c * k * k) ** 0.5)
    wp_ = pack_conv_weight(w, dtype)
    out = torch.zeros(n, hp, wp, cout, dtype=dtype, device="cuda")
    _lib.lib().hoig_set_halo_variant(variant)
    try:
        ops.conv2d_halo([(x.cuda(), wp_.cuda(), out)], k, k, cout)
        torch.cuda.synchronize()
    finally:
        _lib.lib().hoig_set_halo_variant(0)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(dtype).float(), None, padding=k // 2).permute(0, 2, 3, 1)
    r = k // 2
    return out.cpu().float()[:, r:hp - r, r:wp - r], ref[:, r:hp - r, r:wp - r]


@pytest.mark.parametrize("variant", [0, 2], ids=["shifted_descriptors", "box_per_tap"])
@pytest.mark.parametrize("shape", [(2, 20, 20, 64, 5, 128), (1, 36, 36, 128, 5, 128), (3, 17, 23, 64, 3, 64), (1, 40, 40, 192, 7, 16),
                                   (2, 9, 300, 64, 5, 128)],
                         ids=["c64_k5", "c128_k5", "ragged_k3_n64", "c192_k7_n16", "wide_rows"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv2d_halo_interior_exact(shape, dtype, variant):
    n, hp, wp, c, k, cout = shape
    out, ref = _halo_case(n, hp, wp, c, k, cout, dtype, variant)
    rel = ((out - ref).norm() / ref.norm()).item()
    print(f"halo {shape} {dtype} variant {variant}: maxabs {(out - ref).abs().max().item():.3e} relL2 {rel:.3e}")
    assert rel <= (4e-3 if dtype == torch.bfloat16 else 6e-4)
    assert (out - ref).abs().max().item() <= TOL[dtype] * 4


def test_conv2d_halo_two_segments_one_launch():
    g = torch.Generator().manual_seed(1)
    dtype, k, cout = torch.bfloat16, 5, 128
    xs = [torch.randn(2, 20, 20, 64, generator=g).to(dtype), torch.randn(2, 24, 24, 128, generator=g).to(dtype)]
    ws = [torch.randn(cout, x.shape[3], k, k, generator=g) * 0.02 for x in xs]
    outs = [torch.zeros(*x.shape[:3], cout, dtype=dtype, device="cuda") for x in xs]
    ops.conv2d_halo([(x.cuda(), pack_conv_weight(w, dtype).cuda(), o) for x, w, o in zip(xs, ws, outs)], k, k, cout)
    torch.cuda.synchronize()
    for x, w, o in zip(xs, ws, outs):
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(dtype).float(), None, padding=2).permute(0, 2, 3, 1)[:, 2:-2, 2:-2]
        got = o.cpu().float()[:, 2:-2, 2:-2]
        assert ((got - ref).norm() / ref.norm()).item() <= 4e-3


def test_conv2d_halo_rejects_bad_arguments():
    x = torch.zeros(1, 8, 8, 32, dtype=torch.bfloat16, device="cuda")      # C not a multiple of 64
    w = torch.zeros(128, 25 * 32, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.conv2d_halo([(x, w, torch.zeros(1, 8, 8, 128, dtype=torch.bfloat16, device="cuda"))], 5, 5, 128)
    xf = torch.zeros(1, 8, 8, 64, device="cuda")
    with pytest.raises(RuntimeError, match="bf16 or fp16"):
        ops.conv2d_halo([(xf, torch.zeros(128, 25 * 64, device="cuda"), torch.zeros(1, 8, 8, 128, device="cuda"))], 5, 5, 128)


def _attn_inputs(n, h, c, k, seed):
    g = torch.Generator().manual_seed(seed)
    src, tgt = torch.randn(n, h, h, c, generator=g), torch.randn(n, h, h, c, generator=g)
    flow = torch.randn(n, h, h, 2, generator=g) * 3.0
    flow[0, 0, 0] = torch.tensor([-30.0, 25.0]); flow[n - 1, h - 1, h - 2] = torch.tensor([40.0, -0.5])
    flow[0, 1, 1] = torch.tensor([1.0, -2.0])       # integer flow: zero fractions
    hid = 128
    w0 = torch.randn(hid, 2 * c, k, k, generator=g) * (1.0 / (2 * c * k * k) ** 0.5)
    b1, w2, b2 = torch.randn(hid, generator=g) * 0.1, torch.randn(k * k, hid, generator=g) * 0.3, torch.randn(k * k, generator=g) * 0.1
    return src, tgt, flow, w0, b1, w2, b2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("k", [5, 3])
def test_attn_combine_vs_contract(dtype, k):
    n, h, c, hid = 2, 14, 64, 128
    src, tgt, flow, w0, b1, w2, b2 = _attn_inputs(n, h, c, k, 3)
    r = k // 2
    g = torch.Generator().manual_seed(9)
    gt = torch.randn(n, h + 2 * r, h + 2 * r, hid, generator=g).to(dtype)
    gs = torch.randn(n, h + 4 * r, h + 4 * r, hid, generator=g).to(dtype)
    src, tgt = src.to(dtype), tgt.to(dtype)
    out = ops.attn_combine(gt.cuda(), gs.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), src.cuda(), flow.cuda(), tgt.cuda(),
                           torch.empty(n, h, h, c, dtype=dtype, device="cuda"), k)
    ref = emu_ops.attn_combine(gt, gs, b1, w2, b2, src, flow, tgt, torch.empty(n, h, h, c), k)
    d = (out.cpu().float() - ref).abs().max().item()
    print(f"attn_combine {dtype} k{k}: maxabs {d:.3e}")
    assert d <= TOL[dtype] * 2.5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_commuted_attention_block_vs_reference_formulation(dtype):
    """pad + halo conv + combine against extract_attn.py:19-28 computed by the oracle's BlockExtractor restatement."""
    n, h, c, k, hid = 2, 16, 128, 5, 128
    src, tgt, flow, w0, b1, w2, b2 = _attn_inputs(n, h, c, k, 7)
    src, tgt = src.to(dtype), tgt.to(dtype)
    w0 = w0.to(dtype).float()
    # reference formulation (fp32 math on the same 16-bit inputs)
    sn, tn, fl = src.float().permute(0, 3, 1, 2), tgt.float().permute(0, 3, 1, 2), flow.permute(0, 3, 1, 2)
    bs, bt = gr.block_extract(sn, fl, k), gr.block_extract(tn, torch.zeros_like(fl), k)
    hidden = F.leaky_relu(F.conv2d(torch.cat([bt, bs], 1), w0, b1, stride=k), 0.01)
    a = F.softmax(F.conv2d(hidden, w2.view(k * k, hid, 1, 1), b2), 1)
    ref = (tn + F.avg_pool2d(gr.local_attn_reshape(a, k) * bs, k, k)).permute(0, 2, 3, 1)
    # CUDA path
    r = k // 2
    dev = dict(dtype=dtype, device="cuda")
    tpad = ops.replicate_pad(tgt.cuda(), torch.empty(n, h + 2 * r, h + 2 * r, c, **dev), r)
    spad = ops.replicate_pad(src.cuda(), torch.empty(n, h + 4 * r, h + 4 * r, c, **dev), 2 * r)
    wt, ws = pack_conv_weight(w0[:, :c], dtype).cuda(), pack_conv_weight(w0[:, c:], dtype).cuda()
    gt, gs = ops.conv2d_halo([(tpad, wt, torch.empty(n, h + 2 * r, h + 2 * r, hid, **dev)),
                              (spad, ws, torch.empty(n, h + 4 * r, h + 4 * r, hid, **dev))], k, k, hid)
    out = ops.attn_combine(gt, gs, b1.cuda(), w2.cuda(), b2.cuda(), src.cuda(), flow.cuda(), tgt.cuda(),
                           torch.empty(n, h, h, c, **dev), k)
    torch.cuda.synchronize()
    d = out.cpu().float() - ref
    rel = (d.norm() / ref.norm()).item()
    print(f"commuted attention block {dtype}: maxabs {d.abs().max().item():.3e} relL2 {rel:.3e}")
    assert rel <= (5e-3 if dtype == torch.bfloat16 else 8e-4)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("flows", ["hogan_like", "large_random", "mixed"])
def test_attn_combine_tensor_core_path_vs_gather_path_and_contract(dtype, flows):
    """attn_combine_tc_kernel (8x8-pixel tiles, source window staged in shared memory, patch sum as an mma.sync GEMM) against the
    per-pixel gather kernel and the op contract.  'hogan_like': normalised coordinates used as pixel offsets (|flow| <= 3, quirk Q1) --
    (almost) every tile takes the staged path; 'large_random': windows exceed the capacity -- every tile takes the in-kernel gather path;
    'mixed': both within one launch."""
    import hoig_b200._lib as L
    n, h, c, k, hid = 2, 32, 128, 5, 128
    src, tgt, flow, w0, b1, w2, b2 = _attn_inputs(n, h, c, k, 11)
    g = torch.Generator().manual_seed(13)
    # the generator's flows: T - identity grid with T in [-1, 1] or the -2 sentinel, smooth except at mask borders (quirk Q1)
    ramp = torch.arange(-1.0, 1.0, 2.0 / h)
    small = torch.stack([-2.0 - ramp[None, :].expand(h, h), 0.6 * torch.sin(torch.arange(h) / 5.0)[:, None].expand(h, h) - ramp[:, None]], -1)
    small = small[None].repeat(n, 1, 1, 1).clone()
    small[:, 10:20, 12:24] += torch.rand(n, 10, 12, 2, generator=g) * 1.5          # a 'valid correspondence' patch with per-pixel jitter
    if flows == "hogan_like":
        flow = small
    elif flows == "mixed":
        flow = torch.where((torch.arange(h)[None, :, None, None] < h // 2), small, flow)
    gt = torch.randn(n, h + 4, h + 4, hid, generator=g).to(dtype)
    gs = torch.randn(n, h + 8, h + 8, hid, generator=g).to(dtype)
    src, tgt = src.to(dtype), tgt.to(dtype)
    outs = []
    try:
        for mode in (0, 1):
            L.lib().hoig_set_attn_tc_mode(mode)
            out = ops.attn_combine(gt.cuda(), gs.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), src.cuda(), flow.cuda(), tgt.cuda(),
                                   torch.empty(n, h, h, c, dtype=dtype, device="cuda"), k)
            torch.cuda.synchronize()
            outs.append(out.cpu().float())
    finally:
        L.lib().hoig_set_attn_tc_mode(1)
    ref = emu_ops.attn_combine(gt, gs, b1, w2, b2, src, flow, tgt, torch.empty(n, h, h, c), k)
    tol = TOL[dtype] * 2.5
    d_ref = (outs[1] - ref).abs().max().item()
    d_modes = (outs[1] - outs[0]).abs().max().item()
    print(f"attn_combine tc {dtype} {flows}: vs contract {d_ref:.3e}, vs gather path {d_modes:.3e}")
    assert d_ref <= tol and d_modes <= tol
    if flows == "large_random":
        assert torch.equal(outs[0], outs[1])          # same arithmetic on the gather path


def test_attn_combine_tc_in_place_and_channel_slices():
    """The generator calls it in place (out = tgt = tsf) on full tensors; also exercise ld > C views."""
    n, h, c, k, hid, dtype = 1, 16, 64, 5, 128, torch.float16
    src, tgt, flow, w0, b1, w2, b2 = _attn_inputs(n, h, c, k, 21)
    flow = torch.rand(n, h, h, 2, generator=torch.Generator().manual_seed(2)) * 5.0 - 3.0
    g = torch.Generator().manual_seed(5)
    gt = torch.randn(n, h + 4, h + 4, hid, generator=g).to(dtype)
    gs = torch.randn(n, h + 8, h + 8, hid, generator=g).to(dtype)
    sbuf = torch.randn(n, h, h, c + 64, generator=g).to(dtype).cuda()
    tbuf = torch.randn(n, h, h, c + 64, generator=g).to(dtype).cuda()
    s_v, t_v = sbuf[..., :c], tbuf[..., 64:]
    ref = emu_ops.attn_combine(gt, gs, b1, w2, b2, s_v.cpu(), flow, t_v.cpu(), torch.empty(n, h, h, c), k)
    ops.attn_combine(gt.cuda(), gs.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), s_v, flow.cuda(), t_v, t_v, k)
    torch.cuda.synchronize()
    assert (t_v.cpu().float() - ref).abs().max().item() <= TOL[dtype] * 2.5
