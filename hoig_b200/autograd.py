"""torch.autograd Functions over the C ABI -- the fp32 training path.

The reference trains through torch.autograd of ``nn.Conv2d`` / ``nn.ConvTranspose2d`` / ``nn.InstanceNorm2d`` plus its own
backward kernels for BlockExtractor and LocalAttnReshape (thirdparty/block_extractor/block_extractor.py:27-45,
thirdparty/local_attn_reshape/local_attn_reshape.py).  Here:

* forward of every convolution = ``hoig_conv2d`` (fp32 implicit-GEMM kernel, NHWC);
* data gradient = another ``hoig_conv2d`` launch (flipped / transposed weights; stride-2 <-> transposed-conv duality; a strided
  conv whose kernel equals its stride is a 1x1 GEMM plus a pixel shuffle);
* weight gradient = ``hoig_conv2d_wgrad_f32``;  InstanceNorm backward = ``hoig_instnorm_backward_f32``;
* BlockExtractor / LocalAttnReshape forward + backward = the kernels of boundary B2.

Tensors at this level are LOGICAL NCHW fp32 (the reference's interface); physically they are NHWC buffers (channels-last views),
so nothing is transposed between ops.  Elementwise glue (ReLU, adds, concatenations, softmax, pooling) is left to torch.
There is no CPU path: every Function raises on non-CUDA tensors via the C-ABI wrappers.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from .packing import ceil_to, pack_conv_weight


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) with any strides -> contiguous NHWC fp32 (B,H,W,Cp), channels zero-padded to a multiple of 8."""
    B, C, H, W = x.shape
    cp = ceil_to(C, 8)
    xh = x.permute(0, 2, 3, 1)
    if x.dtype == torch.float32 and cp == C and xh.is_contiguous():
        return xh
    out = torch.zeros(B, H, W, cp, dtype=torch.float32, device=x.device)
    out[..., :C] = xh
    return out


def _nchw(y: torch.Tensor, c: int) -> torch.Tensor:
    """NHWC buffer -> logical NCHW view of its first ``c`` channels (no copy)."""
    return y[..., :c].permute(0, 3, 1, 2)


def _conv(xh, w_oihw, out_hw, *, stride=1, pad=(0, 0), transposed=False, bias=None):
    """One hoig_conv2d launch on NHWC fp32.  ``w_oihw``: (Cout,Cin,KH,KW), or the ConvTranspose layout (Cin,Cout,3,3) when
    ``transposed``; input channels are matched to xh's padded channel count by zero columns in the packed matrix."""
    B = xh.shape[0]
    if transposed:
        cin, cout, kh, kw = w_oihw.shape
    else:
        cout, cin, kh, kw = w_oihw.shape
    cp = xh.shape[3]
    if cin != cp:                                # the activations carry zero-padded channels: pad the weight's input dim too
        wpad = torch.zeros((cp, cout, kh, kw) if transposed else (cout, cp, kh, kw), dtype=torch.float32, device=w_oihw.device)
        if transposed:
            wpad[:cin] = w_oihw
        else:
            wpad[:, :cin] = w_oihw
        w_oihw = wpad
    wp = pack_conv_weight(w_oihw, torch.float32, transposed=transposed)
    out = torch.empty(B, out_hw[0], out_hw[1], ceil_to(cout, 8), dtype=torch.float32, device=xh.device)
    if out.shape[3] != cout:
        out.zero_()
    ops.conv2d(xh, wp, out, kh=kh, kw=kw, stride=stride, pad=pad[0], pad_w=pad[1],
               mode=ops.CONV_TRANSPOSED if transposed else ops.CONV, bias=bias, cout=cout)
    return out


def _wgrad(xh, gh, cout, cin, kh, kw, stride, pad):
    """dW (cout, cin, kh, kw) = sum over pixels of g (x) x  (hoig_conv2d_wgrad_f32); xh / gh NHWC fp32, channel-padded."""
    B, H, W, cp = xh.shape
    _, OH, OW, gp = gh.shape
    dw = torch.zeros(cout, kh, kw, cp, dtype=torch.float32, device=xh.device)
    _lib.check(_lib.lib().hoig_conv2d_wgrad_f32(xh.data_ptr(), xh.stride(2), gh.data_ptr(), gh.stride(2), dw.data_ptr(), B, H, W, cp,
                                                OH, OW, cout, kh, kw, stride, pad[0], pad[1],
                                                torch.cuda.current_stream().cuda_stream), "conv2d_wgrad_f32")
    return dw[..., :cin].permute(0, 3, 1, 2)


class Conv2dFn(torch.autograd.Function):
    """``F.conv2d(x, w, b, stride, padding)`` with square stride and (pad_h, pad_w) zero padding."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad):
        B, C, H, W = x.shape
        cout, cin, kh, kw = w.shape
        oh, ow = (H + 2 * pad[0] - kh) // stride + 1, (W + 2 * pad[1] - kw) // stride + 1
        xh = _nhwc(x)
        out = _conv(xh, w.detach(), (oh, ow), stride=stride, pad=pad, bias=None if b is None else b.detach().float().contiguous())
        ctx.save_for_backward(xh, w)
        ctx.geom = (C, stride, pad, b is not None)
        return _nchw(out, cout)

    @staticmethod
    def backward(ctx, g):
        xh, w = ctx.saved_tensors
        C, stride, pad, has_bias = ctx.geom
        cout, cin, kh, kw = w.shape
        B, H, W, _ = xh.shape
        gh = _nhwc(g)
        oh, ow = gh.shape[1], gh.shape[2]
        gx = gw = gb = None
        wd = w.detach()
        if ctx.needs_input_grad[0]:
            if stride == 1:
                # full correlation with the flipped kernel, channels swapped
                dxh = _conv(gh, wd.flip(2, 3).transpose(0, 1), (H, W), pad=(kh - 1 - pad[0], kw - 1 - pad[1]))
            elif stride == 2 and kh == 3 and kw == 3 and pad == (1, 1) and H == 2 * oh and W == 2 * ow:
                # the data gradient of a k3 s2 p1 conv IS the k3 s2 p1 op1 transposed conv with the same weight tensor
                dxh = _conv(gh, wd, (H, W), stride=2, pad=(1, 1), transposed=True)
            elif stride == kh and stride == kw and pad == (0, 0) and H == stride * oh and W == stride * ow:
                # non-overlapping windows (the k5 s5 attention conv): a 1x1 GEMM to k*k*Cin channels + pixel shuffle
                w1 = wd.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout, 1, 1)
                t = _conv(gh, w1, (oh, ow))[..., :kh * kw * cin]
                dxh = t.reshape(B, oh, ow, kh, kw, cin).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, cin)
            else:
                # generic strided conv: zero-insert the gradient, then the stride-1 rule
                zh, zw = (oh - 1) * stride + 1, (ow - 1) * stride + 1
                gz = torch.zeros(B, zh, zw, gh.shape[3], dtype=torch.float32, device=g.device)
                gz[:, ::stride, ::stride] = gh
                hh, ww = zh + kh - 1 - 2 * pad[0], zw + kw - 1 - 2 * pad[1]
                dxh = _conv(gz, wd.flip(2, 3).transpose(0, 1), (hh, ww), pad=(kh - 1 - pad[0], kw - 1 - pad[1]))
                if (hh, ww) != (H, W):        # rows / columns the strided window never reached
                    full = torch.zeros(B, H, W, dxh.shape[3], dtype=torch.float32, device=g.device)
                    full[:, :hh, :ww] = dxh
                    dxh = full
            gx = _nchw(dxh, C)
        if ctx.needs_input_grad[1]:
            gw = _wgrad(xh, gh, cout, cin, kh, kw, stride, pad)
        if has_bias and ctx.needs_input_grad[2]:
            gb = g.sum((0, 2, 3))
        return gx, gw, gb, None, None


class ConvTranspose2dFn(torch.autograd.Function):
    """``F.conv_transpose2d(x, w, None, stride=2, padding=1, output_padding=1)`` with a 3x3 kernel (generator.py:118,201)."""

    @staticmethod
    def forward(ctx, x, w):
        B, C, H, W = x.shape
        cin, cout, kh, kw = w.shape
        xh = _nhwc(x)
        out = _conv(xh, w.detach(), (2 * H, 2 * W), stride=2, pad=(1, 1), transposed=True)
        ctx.save_for_backward(xh, w)
        ctx.c = C
        return _nchw(out, cout)

    @staticmethod
    def backward(ctx, g):
        xh, w = ctx.saved_tensors
        cin, cout, kh, kw = w.shape
        B, H, W, _ = xh.shape
        gh = _nhwc(g)
        gx = gw = None
        if ctx.needs_input_grad[0]:            # the k3 s2 p1 conv whose weight tensor (out=Cin, in=Cout) is w itself
            gx = _nchw(_conv(gh, w.detach(), (H, W), stride=2, pad=(1, 1)), ctx.c)
        if ctx.needs_input_grad[1]:            # roles swapped: dW[ci,co,r,s] = sum x[ci] at (y,x) * g[co] at (2y+r-1, 2x+s-1)
            gw = _wgrad(gh, xh, cin, cout, kh, kw, 2, (1, 1))
        return gx, gw


class InstanceNormFn(torch.autograd.Function):
    """``F.instance_norm(x, weight=gamma, bias=beta, eps=eps)`` (biased variance, no running stats)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        B, C, H, W = x.shape
        if C % 8:
            raise ValueError("InstanceNormFn: channel count must be a multiple of 8")
        xh = _nhwc(x)
        stats = torch.zeros(B * C * 2, dtype=torch.float64, device=x.device)
        ops.plane_stats(xh, stats)
        out = torch.empty_like(xh)
        gm = None if gamma is None else gamma.detach().float().contiguous()
        bt = None if beta is None else beta.detach().float().contiguous()
        ops.instnorm_apply(xh, stats, out, gamma=gm, beta=bt, eps=eps)
        ctx.save_for_backward(xh, stats, gm)
        ctx.eps, ctx.affine = eps, gamma is not None
        return _nchw(out, C)

    @staticmethod
    def backward(ctx, g):
        xh, stats, gm = ctx.saved_tensors
        B, H, W, C = xh.shape
        gh = _nhwc(g)
        dx = torch.empty_like(xh)
        scratch = torch.zeros(B * C * 2, dtype=torch.float64, device=g.device)
        dgam = torch.zeros(C, dtype=torch.float32, device=g.device) if ctx.affine else None
        dbet = torch.zeros(C, dtype=torch.float32, device=g.device) if ctx.affine else None
        _lib.check(_lib.lib().hoig_instnorm_backward_f32(
            xh.data_ptr(), xh.stride(2), gh.data_ptr(), gh.stride(2), stats.data_ptr(), gm.data_ptr() if gm is not None else None,
            dx.data_ptr(), dx.stride(2), scratch.data_ptr(), dgam.data_ptr() if dgam is not None else None,
            dbet.data_ptr() if dbet is not None else None, B, H * W, C, ctypes.c_float(ctx.eps),
            torch.cuda.current_stream().cuda_stream), "instnorm_backward_f32")
        return _nchw(dx, C), dgam, dbet, None


class BlockExtractFn(torch.autograd.Function):
    """BlockExtractor (thirdparty/block_extractor/block_extractor.py:12-45): forward and backward on the boundary-B2 kernels."""

    @staticmethod
    def forward(ctx, source, flow, k):
        source, flow = source.contiguous().float(), flow.contiguous().float()
        out = torch.zeros(source.shape[0], source.shape[1], k * flow.shape[2], k * flow.shape[3], device=source.device)
        ops.block_extract(source, flow, out, k)
        ctx.save_for_backward(source, flow)
        ctx.k = k
        return out

    @staticmethod
    def backward(ctx, g):
        source, flow = ctx.saved_tensors
        gs, gf = torch.zeros_like(source), torch.zeros_like(flow)
        ops.block_extract_backward(source, flow, g.contiguous().float(), gs, gf, ctx.k)
        return gs, gf, None


class LocalAttnReshapeFn(torch.autograd.Function):
    """LocalAttnReshape (thirdparty/local_attn_reshape/local_attn_reshape.py): (B,k*k,H,W) -> (B,1,kH,kW)."""

    @staticmethod
    def forward(ctx, x, k):
        x = x.contiguous().float()
        out = torch.zeros(x.shape[0], 1, k * x.shape[2], k * x.shape[3], device=x.device)
        ops.local_attn_reshape(x, out, k)
        ctx.k, ctx.shape = k, x.shape
        return out

    @staticmethod
    def backward(ctx, g):
        gi = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        ops.local_attn_reshape_backward(g.contiguous().float(), gi, ctx.k)
        return gi, None


def conv2d(x, w, b=None, stride=1, padding=0):
    pad = (padding, padding) if isinstance(padding, int) else tuple(padding)
    return Conv2dFn.apply(x, w, b, stride, pad)


def conv_transpose2d(x, w):
    return ConvTranspose2dFn.apply(x, w)


def instance_norm(x, gamma=None, beta=None, eps=1e-5):
    return InstanceNormFn.apply(x, gamma, beta, eps)


def block_extract(source, flow, k):
    return BlockExtractFn.apply(source, flow, k)


def local_attn_reshape(x, k):
    return LocalAttnReshapeFn.apply(x, k)
