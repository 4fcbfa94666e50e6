"""Torch-CPU emulation of the ``hoig_b200.ops`` contracts -- TEST INFRASTRUCTURE ONLY.

Lets the host-side schedule and weight packing of ``GeneratorB200`` be checked
against the oracle in a container without a GPU.  Each function implements the
documented contract of the C-ABI op of the same name (include/hoig_b200.h) with
plain torch ops; it is also the executable specification the CUDA kernels are
tested against on the GPU box.  Never imported by the product.
"""
import torch
import torch.nn.functional as F

from hoig_b200 import ops as real_ops
from hoig_b200.packing import ceil_to
from oracle import generator_ref as gr

ACT = {0: lambda v: v, 1: F.relu, 2: lambda v: F.leaky_relu(v, 0.01), 3: torch.tanh, 4: torch.sigmoid}


def _q(t, like):
    """Round through the storage dtype (bf16 path) so emulation mirrors kernel rounding points."""
    return t.to(like.dtype)


def unpack_weight(wp, cout, kh, kw, cin_p):
    w = wp.float()[:cout, : kh * kw * cin_p].reshape(cout, kh, kw, cin_p)
    return w.permute(0, 3, 1, 2).contiguous()  # OIHW over padded input channels


def unpack_transposed(wp, cout, kh, kw, cin_p, pad):
    """Inverse of pack_conv_weight(transposed=True): [4*Cout][4*Cin_p] parity/tap blocks -> (Cout, Cin_p, KH, KW)."""
    w = torch.zeros(cout, cin_p, kh, kw)
    for a in (0, 1):
        for b in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    r, s_ = a + pad - 2 * dy, b + pad - 2 * dx
                    if 0 <= r < kh and 0 <= s_ < kw:
                        ph, t = a * 2 + (b ^ a), dy * 2 + dx        # packing.parity_block
                        w[:, :, r, s_] = wp.float()[ph * cout:(ph + 1) * cout, t * cin_p:(t + 1) * cin_p]
    return w


def conv2d(x0, weight, out, *, kh, kw, stride=1, pad=0, pad_w=None, mode=0, x1=None, bias=None, act=0, residual=None, stats=None,
           flow=None, cout=None, simt=False, act_table=None, spade_x=None, spade_stats=None, eps=1e-5):
    x = x0 if x1 is None else torch.cat([x0, x1], 3)
    cin_p = x.shape[3]
    cout = out.shape[3] if cout is None else cout
    xn = x.float().permute(0, 3, 1, 2)
    if mode == real_ops.CONV_TRANSPOSED:
        w = unpack_transposed(weight, cout, kh, kw, cin_p, pad)
        y = F.conv_transpose2d(xn, w.permute(1, 0, 2, 3), None, stride=stride, padding=pad, output_padding=1)
    else:
        w = unpack_weight(weight, cout, kh, kw, cin_p)
    if mode == real_ops.CONV:
        y = F.conv2d(xn, w, None, stride=stride, padding=(pad, pad if pad_w is None else pad_w))
    elif mode == real_ops.CONV_TRANSPOSED:
        pass
    else:
        c = x0.shape[3]
        tgt, src = xn[:, :c], xn[:, c:]
        fl = flow.permute(0, 3, 1, 2)
        bs = _q(gr.block_extract(src, fl, kh), x0).float()
        bt = gr.block_extract(tgt, torch.zeros_like(fl), kh)
        y = F.conv2d(torch.cat([bt, bs], 1), w, None, stride=kh)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1)
    if spade_x is not None:
        # hoigConvDesc::spade_x: columns are (gamma, beta) in blocks of 8 channels; write relu(norm(x) * (1 + gamma) + beta)
        n, h, w_, c2 = y.shape
        c = c2 // 2
        gb = y.reshape(n, h, w_, c // 8, 2, 8)
        gamma, beta = gb[..., 0, :].reshape(n, h, w_, c), gb[..., 1, :].reshape(n, h, w_, c)
        st = spade_stats.view(n, c, 2)
        mean = st[..., 0] / (h * w_)
        var = (st[..., 1] / (h * w_) - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var.float() + eps)
        xn_ = (spade_x.float() - mean.float().view(n, 1, 1, c)) * rstd.view(n, 1, 1, c)
        res_ = torch.relu(xn_ * (1 + gamma) + beta)
        out[..., :c] = _q(res_, out)
        return out
    if residual is not None:
        y = y + residual[..., :cout].float()
    if act_table is not None:
        y = torch.stack([ACT[int(a)](y[..., i]) for i, a in enumerate(act_table.tolist())], -1)
    else:
        y = ACT[act](y)
    if stats is not None:
        yf = y.double()      # statistics of the fp32 values, before the 16-bit store rounding (conv_umma.cu epilogue)
        st = torch.stack([yf.sum((1, 2)), (yf * yf).sum((1, 2))], 2)  # (N, C, 2)
        stats += st.reshape(-1)
    out[..., :cout] = _q(y, out)
    return out


def nchw_to_nhwc(x, out):
    out.zero_()
    out[..., : x.shape[1]] = x.permute(0, 2, 3, 1).to(out.dtype)
    return out


def nhwc_to_nchw(x, channels):
    return x[..., :channels].float().permute(0, 3, 1, 2).contiguous()


def seg_resize(seg, out):
    s = F.interpolate(seg, size=out.shape[1:3], mode="nearest")
    return nchw_to_nhwc(s, out)


def seg_unfold3(seg, out):
    b, c = seg.shape[:2]
    _, ho, wo, kpad = out.shape
    r = F.interpolate(seg, size=(ho, wo), mode="nearest")
    cols = F.unfold(r, 3, padding=1).reshape(b, c, 9, ho, wo).permute(0, 3, 4, 2, 1).reshape(b, ho, wo, 9 * c)   # (t, c) order
    out.zero_()
    out[..., : 9 * c] = cols.to(out.dtype)
    return out


def plane_stats(x, stats):
    xf = x.double()
    stats += torch.stack([xf.sum((1, 2)), (xf * xf).sum((1, 2))], 2).reshape(-1)
    return stats


def instnorm_apply(x, stats, out, *, gamma=None, beta=None, gb=None, residual=None, relu=False, eps=1e-5):
    n, h, w, c = x.shape
    st = stats.reshape(n, c, 2)
    mean = st[..., 0] / (h * w)
    var = (st[..., 1] / (h * w) - mean * mean).clamp_min(0)
    rstd = (1.0 / torch.sqrt(var + eps)).float()
    y = (x.float() - mean.float()[:, None, None, :]) * rstd[:, None, None, :]
    if gamma is not None:
        y = y * gamma + beta
    if gb is not None:
        y = y * (1 + gb[..., :c].float()) + gb[..., c:2 * c].float()
    if residual is not None:
        y = y + residual.float()
    if relu:
        y = F.relu(y)
    out.copy_(y.to(out.dtype))
    return out


def resize_flow(T, h, subtract_identity=True):
    t = gr.resize_trans(T, h)
    if subtract_identity:
        t = t - gr.identity_grid(h, T.device)
    return t.contiguous()


def attn_unfold(src, tgt, flow, out, k):
    n, h, _, c = src.shape
    fl = flow.permute(0, 3, 1, 2)
    bs = gr.block_extract(src.float().permute(0, 3, 1, 2), fl, k)                       # (N,C,kh,kh)
    bt = gr.block_extract(tgt.float().permute(0, 3, 1, 2), torch.zeros_like(fl), k)

    def taps(b):   # -> (N,h,h,k*k,C): tap t = (ky,kx) of pixel (y,x) sits at block position (k*y+ky, k*x+kx)
        return b.reshape(n, c, h, k, h, k).permute(0, 2, 4, 3, 5, 1).reshape(n, h, h, k * k, c)

    u = torch.cat([taps(bt), taps(bs)], 4).reshape(n, h, h, k * k * 2 * c)
    out.copy_(u.to(out.dtype))
    return out


def attn_finish(hidden, w2, b2, src, flow, tgt, out, k, unfold=None):
    n, h, _, c = src.shape
    logits = hidden.float() @ w2.t() + b2                      # (N,h,h,k*k)
    a = F.softmax(logits, 3).permute(0, 3, 1, 2)
    if unfold is not None:
        bsu = unfold.float().reshape(n, h, h, k * k, 2 * c)[..., c:]                   # (N,h,h,kk,C)
        res = (a.permute(0, 2, 3, 1)[..., None] * bsu).sum(3) / (k * k)
        out.copy_((tgt.float() + res).to(out.dtype))
        return out
    bs = gr.block_extract(src.float().permute(0, 3, 1, 2), flow.permute(0, 3, 1, 2), k)
    res = F.avg_pool2d(gr.local_attn_reshape(a, k) * bs, k, k).permute(0, 2, 3, 1)
    out.copy_((tgt.float() + res).to(out.dtype))
    return out


def replicate_pad(x, out, pad):
    y = F.pad(x.float().permute(0, 3, 1, 2), (pad, pad, pad, pad), mode="replicate").permute(0, 2, 3, 1)
    out.copy_(y.to(out.dtype))
    return out


def conv2d_halo(segments, kh, kw, cout):
    """Contract: exact conv at pixels >= k//2 inside the raster; the border ring is don't-care (NaN here, so a
    consumer that reads it fails the parity tests)."""
    outs = []
    for x, w, out in segments:
        c = x.shape[3]
        wt = unpack_weight(w, cout, kh, kw, c)
        y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, padding=(kh // 2, kw // 2)).permute(0, 2, 3, 1).contiguous()
        ry, rx = kh // 2, kw // 2
        y[:, :ry] = float("nan"); y[:, y.shape[1] - ry:] = float("nan")
        y[:, :, :rx] = float("nan"); y[:, :, y.shape[2] - rx:] = float("nan")
        out.copy_(y.to(out.dtype))
        outs.append(out)
    return outs


def attn_combine(gt, gs, b1, w2, b2, src, flow, tgt, out, k):
    """include/hoig_b200.h "local attention, tensor-core formulation", restated with torch indexing."""
    n, h, _, c = src.shape
    r = k // 2
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(h), indexing="ij")
    dx, dy = flow[..., 0] + xs.float(), flow[..., 1] + ys.float()
    fx, fy = torch.floor(dx), torch.floor(dy)
    wx, wy = [1 - (dx - fx), dx - fx], [1 - (dy - fy), dy - fy]
    x0, y0 = fx.clamp(-(k + 2), h + k + 2).long(), fy.clamp(-(k + 2), h + k + 2).long()
    b = torch.arange(n).view(n, 1, 1)
    hid = gt[:, r:r + h, r:r + h].float() + b1
    for qy in range(2):
        for qx in range(2):
            cy, cx = (y0 + qy).clamp(-r, h - 1 + r) + 2 * r, (x0 + qx).clamp(-r, h - 1 + r) + 2 * r
            hid = hid + (wy[qy] * wx[qx])[..., None] * gs[b, cy, cx].float()
    hid = F.leaky_relu(hid, 0.01)
    a = F.softmax(hid @ w2.t() + b2, 3)
    res = torch.zeros(n, h, h, c)
    sf = src.float()
    for ty in range(k):
        for tx in range(k):
            for qy in range(2):
                for qx in range(2):
                    py, px = (y0 - r + ty + qy).clamp(0, h - 1), (x0 - r + tx + qx).clamp(0, h - 1)
                    res += (a[..., ty * k + tx] * wy[qy] * wx[qx])[..., None] * sf[b, py, px]
    out.copy_((tgt.float() + res / (k * k)).to(out.dtype))
    return out


def grid_sample(x, grid, out, tgt=None):
    y = F.grid_sample(x.float().permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    y = y.permute(0, 2, 3, 1)
    if tgt is not None:
        y = y + tgt.float()
    out.copy_(y.to(out.dtype))
    return out


def hunfold_nchw(x, out, k):
    b, c, h, w = x.shape
    xp = F.pad(x, (k // 2, k // 2))
    cols = torch.stack([xp[..., s:s + w] for s in range(k)], 1).reshape(b, k * c, h, w)   # channel s*C + c
    out.zero_()
    out[..., : k * c] = cols.permute(0, 2, 3, 1).to(out.dtype)
    return out


def hfold_nchw(z, groups, k, segments, act_table=None):
    b, h, w, _ = z.shape
    zz = z.float()[..., : k * groups].reshape(b, h, w, k, groups)
    zp = F.pad(zz, (0, 0, 0, 0, k // 2, k // 2))                 # pad W
    y = sum(zp[:, :, s:s + w, s] for s in range(k))               # y[x] = sum_s Z[x + s - k//2, s]
    if act_table is not None:
        y = torch.stack([ACT[int(a)](y[..., i]) for i, a in enumerate(act_table.tolist())], -1)
    y = y.permute(0, 3, 1, 2)
    return [y[:, c0:c0 + n].contiguous() for c0, n in segments]


def composite(img_bg, obj, hand, mask_bg, mask_hand):
    return gr.composite(img_bg, obj, hand, mask_bg, mask_hand)


def install(monkeypatch):
    """Route hoig_b200.generator's op calls to this module (CPU)."""
    import hoig_b200.generator as G

    class _Ops:
        pass

    emu = _Ops()
    for k in dir(real_ops):
        if k.isupper():
            setattr(emu, k, getattr(real_ops, k))
    for name in ("conv2d", "nchw_to_nhwc", "nhwc_to_nchw", "seg_resize", "seg_unfold3", "plane_stats", "instnorm_apply", "resize_flow",
                 "attn_finish", "attn_unfold", "replicate_pad", "conv2d_halo", "attn_combine", "grid_sample", "composite", "hunfold_nchw", "hfold_nchw"):
        setattr(emu, name, globals()[name])
    monkeypatch.setattr(G, "ops", emu)
    return emu
