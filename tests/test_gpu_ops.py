"""GPU parity of the generator's building-block kernels against their op contracts
(tests/emu_ops.py, torch CPU) and the C oracle, called through the C ABI."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import oracle
from hoig_b200 import ops
from hoig_b200.packing import ceil_to, pack_conv_weight

from . import emu_ops

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def _load_ref(name):
    path = os.path.join(REF_DIR, name + ".so")
    if not os.path.exists(path):
        return None
    # the reference's rasterizer hard-codes its pybind module name (rasterize_cuda.cpp:194)
    init = "rasterize" if name == "ref_rasterize_cuda" else name
    spec = importlib.util.spec_from_file_location(init, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _report(tag, got, ref, atol, rtol=0.0):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    lim = atol + rtol * ref.abs()
    worst = (err - lim).max().item()
    rel = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
    idx = np.unravel_index(int(err.argmax()), err.shape)
    print(f"[{tag}] maxabs={err.max().item():.3e} at {idx} got={got[idx].item():.5f} ref={ref[idx].item():.5f} relL2={rel:.3e} "
          f"bad={(err > lim).float().mean().item():.4f}")
    return worst <= 0, rel


def _rand(g, *shape, scale=1.0):
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


# ----------------------------------------------------------------------------- B2 ops
def test_block_extract_vs_oracle_and_reference():
    g = torch.Generator().manual_seed(0)
    src = _rand(g, 2, 6, 14, 10)
    flow = _rand(g, 2, 2, 14, 10, scale=1.8 * 3)     # test_block_extractor.py:74-78 shape, larger flow
    out = ops.block_extract(src.cuda(), flow.cuda(), torch.zeros(2, 6, 70, 50, device="cuda"), 5)
    ref = torch.from_numpy(oracle.block_extract(src.numpy(), flow.numpy(), 5))
    assert torch.equal(out.cpu(), ref), (out.cpu() - ref).abs().max()
    mod = _load_ref("ref_block_extractor_cuda")
    if mod is not None:
        o2 = torch.zeros(2, 6, 70, 50, device="cuda")
        mod.forward(src.cuda(), flow.cuda(), o2, 5)
        torch.cuda.synchronize()
        print("block_extract vs reference kernel: maxabs", (o2 - out).abs().max().item(), "bit-equal", torch.equal(o2, out))
        assert (o2 - out).abs().max().item() <= 1e-6


def test_local_attn_reshape_vs_oracle_and_reference():
    x = torch.rand(4, 9, 14, 10)                      # test_local_attn_reshape.py:66 shape
    out = ops.local_attn_reshape(x.cuda(), torch.zeros(4, 1, 42, 30, device="cuda"), 3)
    assert torch.equal(out.cpu(), torch.from_numpy(oracle.local_attn_reshape(x.numpy(), 3)))
    pat = torch.arange(9.0).view(1, 9, 1, 1).expand(1, 9, 4, 4).contiguous()
    o = ops.local_attn_reshape(pat.cuda(), torch.zeros(1, 1, 12, 12, device="cuda"), 3)
    assert torch.equal(o[0, 0, :3, :3].cpu(), torch.arange(9.0).view(3, 3))     # KAT implied by the reference script
    mod = _load_ref("ref_local_attn_reshape_cuda")
    if mod is not None:
        o2 = torch.zeros(4, 1, 42, 30, device="cuda")
        mod.forward(x.cuda(), o2, 3)
        torch.cuda.synchronize()
        assert torch.equal(o2, out)


def test_block_extract_and_reshape_backward_vs_oracle_and_reference():
    """Row N3 (boundary B2 backward entry points): gradients of BlockExtractor / LocalAttnReshape against the C oracle and the
    reference's own backward kernels compiled from /root/reference (atomic summation order differs: tolerance, not bit-equality)."""
    g = torch.Generator().manual_seed(4)
    B, C, H, W, k = 2, 6, 14, 10, 5
    src, flow = torch.randn(B, C, H, W, generator=g), torch.randn(B, 2, H, W, generator=g) * 3.0
    flow[0, :, 0, 0] = torch.tensor([-40.0, 33.0])                   # taps clamped at the border
    gout = torch.randn(B, C, k * H, k * W, generator=g)
    gs = torch.zeros_like(src).cuda()
    gf = torch.zeros_like(flow).cuda()
    ops.block_extract_backward(src.cuda(), flow.cuda(), gout.cuda(), gs, gf, k)
    rs, rf = oracle.block_extract_backward(src.numpy(), flow.numpy(), gout.numpy(), k)
    assert np.abs(gs.cpu().numpy() - rs).max() <= 1e-4 * max(1.0, np.abs(rs).max())
    assert np.abs(gf.cpu().numpy() - rf).max() <= 1e-4 * max(1.0, np.abs(rf).max())
    mod = _load_ref("ref_block_extractor_cuda")
    if mod is not None:
        g2, f2 = torch.zeros_like(src).cuda(), torch.zeros_like(flow).cuda()
        mod.backward(src.cuda(), flow.cuda(), gout.cuda(), g2, f2, k)
        torch.cuda.synchronize()
        print("block_extract backward vs reference kernel: grad_source maxabs", (g2 - gs).abs().max().item(), "grad_flow maxabs",
              (f2 - gf).abs().max().item())
        assert (g2 - gs).abs().max().item() <= 1e-4 * max(1.0, g2.abs().max().item())
        assert (f2 - gf).abs().max().item() <= 1e-4 * max(1.0, f2.abs().max().item())
    # accumulate semantics: a second call doubles the result
    ops.block_extract_backward(src.cuda(), flow.cuda(), gout.cuda(), gs, gf, k)
    assert np.abs(gs.cpu().numpy() - 2 * rs).max() <= 2e-4 * max(1.0, np.abs(rs).max())
    x = torch.rand(4, 9, 14, 10)
    go = torch.randn(4, 1, 42, 30, generator=g)
    gi = ops.local_attn_reshape_backward(go.cuda(), torch.zeros_like(x).cuda(), 3)
    assert np.array_equal(gi.cpu().numpy(), oracle.local_attn_reshape_backward(go.numpy(), 3))
    mod = _load_ref("ref_local_attn_reshape_cuda")
    if mod is not None:
        gi2 = torch.zeros_like(x).cuda()
        mod.backward(x.cuda(), go.cuda(), gi2, 3)
        torch.cuda.synchronize()
        assert torch.equal(gi2, gi)


def test_compat_autograd_functions_run_forward_and_backward():
    """The reference's autograd Functions (block_extractor.py:6-42, local_attn_reshape.py:6-37) call `forward` / `backward` of the
    extension modules; the stand-ins must support both so `loss.backward()` through the unmodified wrappers works."""
    import hoig_b200.compat as compat
    be, lar, _ = compat.install()
    g = torch.Generator().manual_seed(2)
    src = torch.randn(1, 4, 8, 8, generator=g).cuda()
    flow = (torch.randn(1, 2, 8, 8, generator=g) * 2).cuda()
    out = torch.zeros(1, 4, 24, 24, device="cuda")
    assert be.forward(src, flow, out, 3) == 1
    gs, gf = torch.zeros_like(src), torch.zeros_like(flow)
    assert be.backward(src, flow, torch.ones_like(out), gs, gf, 3) == 1
    # d(sum of all samples)/d(source) sums the bilinear weights landing on each source pixel: total = number of samples
    assert abs(gs.sum().item() - out.numel()) <= 1e-2
    gi = torch.zeros(1, 9, 8, 8, device="cuda")
    assert lar.backward(torch.zeros(1, 9, 8, 8, device="cuda"), torch.ones(1, 1, 24, 24, device="cuda"), gi, 3) == 1
    assert torch.equal(gi, torch.ones_like(gi))


# ------------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_layout_and_norm_ops(dtype):
    g = torch.Generator().manual_seed(1)
    x = _rand(g, 2, 12, 16, 16)
    nhwc = ops.nchw_to_nhwc(x.cuda(), torch.empty(2, 16, 16, 16, dtype=dtype, device="cuda"))
    ref = emu_ops.nchw_to_nhwc(x, torch.empty(2, 16, 16, 16, dtype=dtype))
    assert torch.equal(nhwc.cpu(), ref)
    assert torch.equal(ops.nhwc_to_nchw(nhwc, 12).cpu(), emu_ops.nhwc_to_nchw(ref, 12))
    seg = torch.rand(2, 12, 64, 64, generator=g)
    for h in (32, 16, 8):
        a = ops.seg_resize(seg.cuda(), torch.empty(2, h, h, 16, dtype=dtype, device="cuda"))
        assert torch.equal(a.cpu(), emu_ops.seg_resize(seg, torch.empty(2, h, h, 16, dtype=dtype)))
    # statistics + normalisation on a channel-slice view (ld > C)
    for C, HW in ((16, 8), (64, 32), (512, 8)):
        buf = (_rand(g, 3, HW, HW, 2 * C) * 2 + 0.5).to(dtype)
        xv = buf[..., :C]
        st = ops.plane_stats(xv.cuda(), torch.zeros(3 * C * 2, dtype=torch.float64, device="cuda"))
        st_ref = emu_ops.plane_stats(xv, torch.zeros(3 * C * 2, dtype=torch.float64))
        assert torch.allclose(st.cpu(), st_ref, rtol=1e-5, atol=1e-3)
        gamma, beta = torch.rand(C, generator=g) + 0.5, _rand(g, C)
        gb = _rand(g, 3, HW, HW, 2 * C).to(dtype)
        res = _rand(g, 3, HW, HW, C).to(dtype)
        tol = {torch.float32: 1e-5, torch.bfloat16: 4e-2, torch.float16: 5e-3}[dtype]
        for kw in (dict(gamma=gamma, beta=beta, relu=True), dict(gamma=gamma, beta=beta, residual=res), dict(gb=gb, relu=True)):
            kw_c = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
            out = ops.instnorm_apply(xv.cuda(), st, torch.empty(3, HW, HW, C, dtype=dtype, device="cuda"), **kw_c)
            ref = emu_ops.instnorm_apply(xv, st_ref, torch.empty(3, HW, HW, C, dtype=dtype), **kw)
            ok, _ = _report(f"instnorm C={C} {dtype} {sorted(kw)}", out, ref, tol, tol)
            assert ok


def test_resize_flow_and_warp_ops():
    g = torch.Generator().manual_seed(2)
    T = _rand(g, 2, 64, 64, 2)
    T[torch.rand(2, 64, 64, generator=g) < 0.3] = -2.0
    for h in (32, 16, 8):
        for sub in (True, False):
            a = ops.resize_flow(T.cuda(), h, sub)
            b = emu_ops.resize_flow(T, h, sub)
            assert (a.cpu() - b).abs().max().item() <= 2e-6
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 3e-2)):
        C, h = 32, 16
        src, tgt = _rand(g, 2, h, h, C).to(dtype), _rand(g, 2, h, h, C).to(dtype)
        flow = emu_ops.resize_flow(T, h, True)
        hidden = _rand(g, 2, h, h, 128).to(dtype)
        w2, b2 = _rand(g, 25, 128, scale=0.2), _rand(g, 25)
        out = ops.attn_finish(hidden.cuda(), w2.cuda(), b2.cuda(), src.cuda(), flow.cuda(), tgt.cuda(),
                              torch.empty(2, h, h, C, dtype=dtype, device="cuda"), 5)
        ref = emu_ops.attn_finish(hidden, w2, b2, src, flow, tgt, torch.empty(2, h, h, C, dtype=dtype), 5)
        ok, _ = _report(f"attn_finish {dtype}", out, ref, tol, tol)
        assert ok
        grid = emu_ops.resize_flow(T, h, False)
        out = ops.grid_sample(src.cuda(), grid.cuda(), torch.empty(2, h, h, C, dtype=dtype, device="cuda"), tgt=tgt.cuda())
        ref = emu_ops.grid_sample(src, grid, torch.empty(2, h, h, C, dtype=dtype), tgt=tgt)
        ok, _ = _report(f"grid_sample {dtype}", out, ref, tol, tol)
        assert ok
    imgs = [_rand(g, 2, 3, 8, 8) for _ in range(3)] + [torch.rand(2, 1, 8, 8, generator=g) for _ in range(2)]
    out = ops.composite(*[t.cuda() for t in imgs])
    assert (out.cpu() - emu_ops.composite(*imgs)).abs().max().item() <= 1e-6


def test_fold_and_unfold_ops():
    g = torch.Generator().manual_seed(5)
    x = _rand(g, 2, 8, 16, 24)
    for dtype in (torch.float32, torch.bfloat16, torch.float16):
        a = ops.hunfold_nchw(x.cuda(), torch.empty(2, 16, 24, 64, dtype=dtype, device="cuda"), 7)
        b = emu_ops.hunfold_nchw(x, torch.empty(2, 16, 24, 64, dtype=dtype), 7)
        assert torch.equal(a.cpu(), b)
        z = _rand(g, 2, 16, 24, 56).to(dtype)
        table = torch.tensor([3, 3, 3, 4, 4, 3, 3, 3], dtype=torch.int32)
        segs = [(0, 3), (3, 1), (4, 1), (5, 3)]
        outs = ops.hfold_nchw(z.cuda(), 8, 7, segs, table.cuda())
        refs = emu_ops.hfold_nchw(z, 8, 7, segs, table)
        for o, r in zip(outs, refs):
            assert o.shape == r.shape and (o.cpu() - r).abs().max().item() <= 2e-6
        C, h = 32, 16
        src, tgt = _rand(g, 2, h, h, C).to(dtype), _rand(g, 2, h, h, C).to(dtype)
        flow = _rand(g, 2, h, h, 2, scale=3.0)
        u = ops.attn_unfold(src.cuda(), tgt.cuda(), flow.cuda(), torch.empty(2, h, h, 50 * C, dtype=dtype, device="cuda"), 5)
        ur = emu_ops.attn_unfold(src, tgt, flow, torch.empty(2, h, h, 50 * C, dtype=dtype), 5)
        tol = {torch.float32: 2e-6, torch.bfloat16: 1.6e-2, torch.float16: 2e-3}[dtype]
        assert (u.cpu().float() - ur.float()).abs().max().item() <= tol
        hidden = _rand(g, 2, h, h, 128).to(dtype)
        w2, b2 = _rand(g, 25, 128, scale=0.2), _rand(g, 25)
        o = ops.attn_finish(hidden.cuda(), w2.cuda(), b2.cuda(), src.cuda(), flow.cuda(), tgt.cuda(),
                            torch.empty(2, h, h, C, dtype=dtype, device="cuda"), 5, unfold=u)
        r = emu_ops.attn_finish(hidden, w2, b2, src, flow, tgt, torch.empty(2, h, h, C, dtype=dtype), 5, unfold=ur)
        ok, _ = _report(f"attn_finish(unfold) {dtype}", o, r, {torch.float32: 2e-5, torch.bfloat16: 3e-2, torch.float16: 4e-3}[dtype], 3e-2)
        assert ok


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C,k,cpad", [(9, 7, 64), (3, 7, 24), (8, 3, 24), (16, 5, 80)])
def test_hunfold_row_kernel_matches_contract(dtype, C, k, cpad):
    """Rows that are multiples of 128 pixels take the shared-memory staged kernel."""
    g = torch.Generator().manual_seed(11)
    x = _rand(g, 2, C, 5, 256)
    a = ops.hunfold_nchw(x.cuda(), torch.empty(2, 5, 256, cpad, dtype=dtype, device="cuda"), k)
    b = emu_ops.hunfold_nchw(x, torch.empty(2, 5, 256, cpad, dtype=dtype), k)
    assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("C,hi,ho", [(3, 32, 8), (12, 16, 16), (12, 20, 40), (5, 9, 7), (12, 256, 128), (3, 256, 32), (12, 256, 64)])
def test_seg_unfold3_matches_contract(dtype, C, hi, ho):
    g = torch.Generator().manual_seed(13)
    seg = _rand(g, 2, C, hi, hi)
    kpad = -(-9 * C // 64) * 64
    a = ops.seg_unfold3(seg.cuda(), torch.empty(2, ho, ho, kpad, dtype=dtype, device="cuda"))
    b = emu_ops.seg_unfold3(seg, torch.empty(2, ho, ho, kpad, dtype=dtype))
    assert torch.equal(a.cpu(), b)


def test_fp16_storage_saturates_instead_of_overflowing():
    """Values beyond the fp16 range are stored as +-65504, not inf (layout kernel and conv epilogue share the conversion)."""
    x = torch.tensor([1e6, -1e6, 3.0, 70000.0, -65504.0, 0.5, 1e38, -1e38]).view(1, 8, 1, 1)
    out = ops.nchw_to_nhwc(x.cuda(), torch.empty(1, 1, 1, 8, dtype=torch.float16, device="cuda"))
    got = out.cpu().float().view(-1)
    assert torch.isfinite(got).all()
    assert got.tolist() == [65504.0, -65504.0, 3.0, 65504.0, -65504.0, 0.5, 65504.0, -65504.0]


# --------------------------------------------------------------------------- convolution
CONV_CASES = [
    # name, N, H, Cin, Cout, k, stride, mode, extras
    ("3x3_s1_c64", 2, 16, 64, 64, 3, 1, "conv", dict(stats=True)),
    ("3x3_s1_c128_n512_bias_res", 1, 32, 128, 512, 3, 1, "conv", dict(bias=True, residual=True, stats=True)),
    ("3x3_s1_k4608", 1, 32, 512, 512, 3, 1, "conv", dict(stats=True)),
    ("3x3_s1_w64", 1, 64, 64, 128, 3, 1, "conv", dict(bias=True, act=1)),
    ("3x3_s1_w256_many_tiles", 3, 256, 64, 64, 3, 1, "conv", dict(stats=True)),
    ("3x3_s2", 2, 32, 64, 128, 3, 2, "conv", dict(stats=True)),
    ("3x3_s2_c32_gather_views", 2, 16, 32, 64, 3, 2, "conv", dict(stats=True, bias=True)),
    ("3x3_s2_c256", 1, 64, 256, 512, 3, 2, "conv", dict(stats=True)),
    ("3x3_s2_w256", 1, 256, 64, 128, 3, 2, "conv", dict(stats=True)),        # the 256 -> 128 stride-2 level: full-row tiles over the four parity views
    ("7x7_heads_merged_act_table", 1, 32, 128, 8, 7, 1, "conv", dict(act_table=[3, 3, 3, 4, 4, 3, 3, 3])),
    ("convT_3x3_s2_c512", 1, 32, 512, 256, 3, 2, "convT", dict(stats=True)),
    ("7x7_stem_c8", 2, 32, 8, 64, 7, 1, "conv", dict(stats=True)),
    ("3x3_seg_c16_relu", 2, 16, 16, 128, 3, 1, "conv", dict(bias=True, act=1)),
    ("7x7_head_c64_n3_tanh", 1, 32, 64, 3, 7, 1, "conv", dict(act=3)),
    ("7x7_head_c128_n1_sigmoid", 1, 32, 128, 1, 7, 1, "conv", dict(act=4)),
    ("1x1", 2, 16, 64, 32, 1, 1, "conv", dict(bias=True)),
    ("1x1_k3200_attn_gemm", 1, 16, 3200, 128, 1, 1, "conv", dict(bias=True, act=2)),
    ("7x1_stem_c64", 2, 32, 64, 64, 7, 1, "conv", dict(stats=True, kw=1)),
    ("7x1_heads_c128_n56", 1, 32, 128, 56, 7, 1, "conv", dict(kw=1)),
    ("7x1_stem_c64_w256", 3, 256, 64, 64, 7, 1, "conv", dict(stats=True, kw=1)),
    ("7x1_bg_head_c64_n21", 2, 64, 64, 21, 7, 1, "conv", dict(kw=1)),
    ("3x3_tiny_8x8", 2, 8, 32, 32, 3, 1, "conv", dict(stats=True)),
    ("convT_3x3_s2", 2, 16, 128, 64, 3, 2, "convT", dict(stats=True)),
    ("convT_3x3_s2_small", 1, 8, 32, 16, 3, 2, "convT", dict(stats=True)),
    ("attn_c64", 2, 16, 64, 128, 5, 5, "attn", dict(bias=True, act=2)),
    ("attn_c32", 1, 16, 32, 128, 5, 5, "attn", dict(bias=True, act=2)),
    ("attn_c512", 1, 32, 512, 128, 5, 5, "attn", dict(bias=True, act=2)),
]


def _run_conv(case, dtype, simt=False):
    name, N, H, Cin, Cout, k, stride, mode, ex = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    pad = k // 2
    kw_ = ex.get("kw", k)
    padw = kw_ // 2
    if mode == "attn":
        tgt, src = _rand(g, N, H, H, Cin).to(dtype), _rand(g, N, H, H, Cin).to(dtype)
        flow = _rand(g, N, H, H, 2, scale=3.0)
        w = _rand(g, Cout, 2 * Cin, k, k, scale=0.05)
        OH, m, x0, x1, pad_ = H, ops.CONV_LOCAL_ATTN, tgt, src, 0
    elif mode == "convT":
        x0, x1, flow = _rand(g, N, H, H, Cin).to(dtype), None, None
        w = _rand(g, Cin, Cout, k, k, scale=0.05)
        OH, m, pad_ = 2 * H, ops.CONV_TRANSPOSED, 1
    else:
        # exercise the channel-slice view (ld > C) on the input
        buf = _rand(g, N, H, H, Cin + 8).to(dtype)
        x0, x1, flow = buf[..., :Cin], None, None
        w = _rand(g, Cout, Cin, k, kw_, scale=0.05)
        OH, m, pad_ = (H + 2 * pad - k) // stride + 1, ops.CONV, pad
    wp = pack_conv_weight(w, dtype, transposed=(mode == "convT"))
    bias = _rand(g, Cout) if ex.get("bias") else None
    Cst = ceil_to(Cout, 8)
    res = _rand(g, N, OH, OH, Cst).to(dtype) if ex.get("residual") else None
    kw = dict(kh=k, kw=kw_ if mode == "conv" else k, stride=stride, pad=pad_, mode=m, act=ex.get("act", 0), cout=Cout)
    if mode == "conv":
        kw["pad_w"] = padw
    table = torch.tensor(ex["act_table"], dtype=torch.int32) if ex.get("act_table") else None
    st_ref = torch.zeros(N * Cout * 2, dtype=torch.float64) if ex.get("stats") else None
    ref = emu_ops.conv2d(x0, wp, torch.zeros(N, OH, OH, Cst, dtype=dtype), x1=x1, bias=bias, residual=res, stats=st_ref, flow=flow,
                         act_table=table, **kw)
    st = torch.zeros(N * Cout * 2, dtype=torch.float64, device="cuda") if ex.get("stats") else None
    cu = lambda t: None if t is None else t.cuda()
    out = torch.zeros(N, OH, OH, Cst, dtype=dtype, device="cuda")
    ops.conv2d(cu(x0), wp.cuda(), out, x1=cu(x1), bias=cu(bias), residual=cu(res), stats=st, flow=cu(flow), simt=simt,
               act_table=cu(table), **kw)
    torch.cuda.synchronize()
    return out, ref, st, st_ref


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_fp32_simt(case):
    out, ref, st, st_ref = _run_conv(case, torch.float32)
    Cout = case[4]
    ok, _ = _report("simt f32 " + case[0], out[..., :Cout], ref[..., :Cout], 2e-4, 2e-4)
    assert ok
    if st is not None:
        assert torch.allclose(st.cpu(), st_ref, rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_bf16_simt_cross_check(case):
    out, ref, st, st_ref = _run_conv(case, torch.bfloat16, simt=True)
    Cout = case[4]
    ok, rel = _report("simt bf16 " + case[0], out[..., :Cout], ref[..., :Cout], 2e-2, 2e-2)
    assert ok and rel < 1e-2
