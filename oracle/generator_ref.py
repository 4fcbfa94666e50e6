"""Functional PyTorch restatement of the reference generator -- TEST INFRASTRUCTURE ONLY.

A pure function of a reference-layout ``state_dict`` (425 entries for the
shipped HOv3 config) and the 12 generator inputs.  Runs on CPU in fp32 (or on
any torch device, the maths is device-agnostic) and restates, citing
/root/reference/HOIG_HOv3/models/networks:

* ``generator.py:347-376``  Generator.forward (bg input assembly, two bg passes)
* ``generator.py:379-464``  infer_front (dual-stream schedule, 9 warps)
* ``generator.py:93-135``   ResNetGenerator
* ``generator.py:138-315``  ResUnetGenerator encode / resnets / decode / regress
* ``generator.py:9-90``     ResidualBlock, SPADEResidualBlock, SPADEBlock
* ``spade.py:24-38``        SPADE
* ``extract_attn.py:23-29`` ExtractorAttn.forward, on top of the restated
  BlockExtractor / LocalAttnReshape (thirdparty/*/..._kernel.cu)
* ``generator.py:466-491``  resize_trans / stn / transform (quirks Q1-Q4 kept)

Parity pinning: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference
``Generator`` class (imported from /root/reference with import stubs and this
module's CPU BlockExtractor/LocalAttnReshape standing in for the two CUDA-only
ops) and commits strided samples of its outputs; ``tests/test_oracle_generator``
checks this restatement against them, and against the live reference when
/root/reference is present.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- ops
def block_extract(src: torch.Tensor, flow: torch.Tensor, k: int) -> torch.Tensor:
    """thirdparty/block_extractor/block_extractor_kernel.cu:52-84 as gathers.

    src (B,C,Hs,Ws), flow (B,2,Hf,Wf) in *pixel* units -> (B,C,k*Hf,k*Wf).
    Border handling is index clamp; the four taps are accumulated in the
    kernel's order LT, RT, LB, RB.
    """
    B, C, Hs, Ws = src.shape
    _, _, Hf, Wf = flow.shape
    dev = src.device
    y = torch.arange(k * Hf, device=dev)
    x = torch.arange(k * Wf, device=dev)
    yf, xf = y // k, x // k
    yo = (y % k - k // 2).to(src.dtype)
    xo = (x % k - k // 2).to(src.dtype)
    fy = flow[:, 1][:, yf][:, :, xf] + yo[None, :, None]
    fx = flow[:, 0][:, yf][:, :, xf] + xo[None, None, :]
    dy = fy + yf.to(src.dtype)[None, :, None]
    dx = fx + xf.to(src.dtype)[None, None, :]
    fdx, fdy = torch.floor(dx), torch.floor(dy)
    xL = fdx.long().clamp(0, Ws - 1)
    xR = (fdx + 1).long().clamp(0, Ws - 1)
    yT = fdy.long().clamp(0, Hs - 1)
    yB = (fdy + 1).long().clamp(0, Hs - 1)
    xLp, xRp = 1 - (dx - fdx), dx - fdx
    yTp, yBp = 1 - (dy - fdy), dy - fdy
    flat = src.reshape(B, C, Hs * Ws)

    def tap(yy, xx):
        idx = (yy * Ws + xx).reshape(B, 1, -1).expand(B, C, -1)
        return flat.gather(2, idx).reshape(B, C, k * Hf, k * Wf)

    out = (xLp * yTp)[:, None] * tap(yT, xL)
    out = out + (xRp * yTp)[:, None] * tap(yT, xR)
    out = out + (xLp * yBp)[:, None] * tap(yB, xL)
    out = out + (xRp * yBp)[:, None] * tap(yB, xR)
    return out


def local_attn_reshape(x: torch.Tensor, k: int) -> torch.Tensor:
    """thirdparty/local_attn_reshape/local_attn_reshape_kernel.cu:47-58."""
    return F.pixel_shuffle(x, k)


def _inorm(x, w=None, b=None):
    return F.instance_norm(x, None, None, w, b, True, 0.1, 1e-5)


def spade(sd: SD, p: str, x: torch.Tensor, seg: torch.Tensor) -> torch.Tensor:
    """spade.py:24-38."""
    n = _inorm(x)
    s = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    a = F.relu(F.conv2d(s, sd[p + "mlp_shared.0.weight"], sd[p + "mlp_shared.0.bias"], padding=1))
    gamma = F.conv2d(a, sd[p + "mlp_gamma.weight"], sd[p + "mlp_gamma.bias"], padding=1)
    beta = F.conv2d(a, sd[p + "mlp_beta.weight"], sd[p + "mlp_beta.bias"], padding=1)
    return n * (1 + gamma) + beta


def residual_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """generator.py:9-32 (dim_in == dim_out, identity shortcut)."""
    h = F.conv2d(x, sd[p + "main.0.weight"], None, padding=1)
    h = F.relu(_inorm(h, sd[p + "main.1.weight"], sd[p + "main.1.bias"]))
    h = F.conv2d(h, sd[p + "main.3.weight"], None, padding=1)
    h = _inorm(h, sd[p + "main.4.weight"], sd[p + "main.4.bias"])
    return x + h


def spade_residual_block(sd: SD, p: str, x: torch.Tensor, seg: torch.Tensor) -> torch.Tensor:
    """generator.py:35-71."""
    dx = F.conv2d(F.relu(spade(sd, p + "norm_0.", x, seg)), sd[p + "conv_0.weight"], sd[p + "conv_0.bias"], padding=1)
    dx = F.conv2d(F.relu(spade(sd, p + "norm_1.", dx, seg)), sd[p + "conv_1.weight"], sd[p + "conv_1.bias"], padding=1)
    return x + dx


def _conv_in_relu(sd: SD, p: str, x, stride=1, padding=1, transposed=False):
    """The ubiquitous ``Sequential(conv, InstanceNorm2d(affine), ReLU)``."""
    if transposed:
        h = F.conv_transpose2d(x, sd[p + "0.weight"], None, stride=2, padding=1, output_padding=1)
    else:
        h = F.conv2d(x, sd[p + "0.weight"], None, stride=stride, padding=padding)
    return F.relu(_inorm(h, sd[p + "1.weight"], sd[p + "1.bias"]))


# ------------------------------------------------------------------ sub-networks
def resnet_generator(sd: SD, p: str, x: torch.Tensor, n_down: int, repeat_num: int) -> torch.Tensor:
    """generator.py:93-135 (bg_model); Sequential indices are part of the key layout."""
    i = 0
    h = F.conv2d(x, sd[f"{p}model.{i}.weight"], None, padding=3)
    h = F.relu(_inorm(h, sd[f"{p}model.{i+1}.weight"], sd[f"{p}model.{i+1}.bias"]))
    i += 3
    for _ in range(n_down):
        h = F.conv2d(h, sd[f"{p}model.{i}.weight"], None, stride=2, padding=1)
        h = F.relu(_inorm(h, sd[f"{p}model.{i+1}.weight"], sd[f"{p}model.{i+1}.bias"]))
        i += 3
    for _ in range(repeat_num):
        h = residual_block(sd, f"{p}model.{i}.", h)
        i += 1
    for _ in range(n_down):
        h = F.conv_transpose2d(h, sd[f"{p}model.{i}.weight"], None, stride=2, padding=1, output_padding=1)
        h = F.relu(_inorm(h, sd[f"{p}model.{i+1}.weight"], sd[f"{p}model.{i+1}.bias"]))
        i += 3
    return torch.tanh(F.conv2d(h, sd[f"{p}model.{i}.weight"], None, padding=3))


class _UNet:
    """ResUnetGenerator (generator.py:138-315) as functions over a key prefix."""

    def __init__(self, sd: SD, p: str, n_down: int, repeat_num: int, spade_layers: Sequence[int]):
        self.sd, self.p, self.n_down, self.repeat_num, self.sl = sd, p, n_down, repeat_num, list(spade_layers)

    def stem(self, x):
        return _conv_in_relu(self.sd, self.p + "encoders.0.", x, padding=3)

    def encoder(self, i, x, seg):
        p = f"{self.p}encoders.{i}."
        if self.sl[0]:  # SPADEBlock, generator.py:74-90
            h = F.conv2d(x, self.sd[p + "conv.weight"], None, stride=2, padding=1)
            return F.relu(spade(self.sd, p + "norm.", h, seg))
        return _conv_in_relu(self.sd, p, x, stride=2)

    def resnet(self, i, x, seg):
        p = f"{self.p}resnets.{i}."
        use_spade = self.sl[1] if i < self.repeat_num // 2 else self.sl[2]
        return spade_residual_block(self.sd, p, x, seg) if use_spade else residual_block(self.sd, p, x)

    def decode(self, x, enc_outs, seg):
        d = x
        for i in range(self.n_down):
            p = f"{self.p}decoders.{i}."
            if self.sl[3]:
                h = F.conv_transpose2d(d, self.sd[p + "conv.weight"], None, stride=2, padding=1, output_padding=1)
                d = F.relu(spade(self.sd, p + "norm.", h, seg))
            else:
                d = _conv_in_relu(self.sd, p, d, transposed=True)
            d = torch.cat([enc_outs[self.n_down - 1 - i], d], 1)
            d = _conv_in_relu(self.sd, f"{self.p}skippers.{i}.", d)
        return d

    def forward(self, x, seg):
        h = self.stem(x)
        outs = [h]
        for i in range(1, self.n_down + 1):
            h = self.encoder(i, h, seg)
            outs.append(h)
        for i in range(self.repeat_num):
            h = self.resnet(i, h, seg)
        return self.decode(h, outs, seg)

    def head(self, name, x, act):
        return act(F.conv2d(x, self.sd[f"{self.p}{name}.0.weight"], None, padding=3))


# ------------------------------------------------------------------------ warps
def resize_trans(T: torch.Tensor, h: int) -> torch.Tensor:
    """generator.py:466-473; size=(h,h) and the -2 sentinels are blended (Q3)."""
    t = F.interpolate(T.permute(0, 3, 1, 2), size=(h, h), mode="bilinear", align_corners=True)
    return t.permute(0, 2, 3, 1)


def identity_grid(h: int, device) -> torch.Tensor:
    """generator.py:484-487: 'ij' meshgrid, channel 0 varies along rows (Q2)."""
    a = torch.arange(start=-1.0, end=1.0, step=2.0 / h)
    xx, yy = torch.meshgrid(a, a, indexing="ij")
    return torch.stack([xx, yy], 2)[None].to(device)


def attn_warp(sd: SD, p: str, src, tgt, flow, k: int = 5):
    """extract_attn.py:23-29."""
    bs = block_extract(src, flow, k)
    bt = block_extract(tgt, torch.zeros_like(flow), k)
    h = F.conv2d(torch.cat((bt, bs), 1), sd[p + "fully_connect_layer.0.weight"],
                 sd[p + "fully_connect_layer.0.bias"], stride=k)
    h = F.leaky_relu(h, 0.01)
    a = F.softmax(F.conv2d(h, sd[p + "fully_connect_layer.2.weight"], sd[p + "fully_connect_layer.2.bias"]), 1)
    a = local_attn_reshape(a, k)
    return F.avg_pool2d(a * bs, k, k)


def transform(sd: SD, x, T, y=None, attn_prefix: Optional[str] = None):
    """generator.py:480-491."""
    Ts = resize_trans(T, x.shape[2])
    if attn_prefix is not None:
        flow = (Ts - identity_grid(x.shape[2], x.device)).permute(0, 3, 1, 2)  # Q1: normalised units used as pixels
        return attn_warp(sd, attn_prefix, x, y, flow)
    return F.grid_sample(x, Ts, mode="bilinear", padding_mode="zeros", align_corners=False)  # Q4


# -------------------------------------------------------------------- generator
def generator_forward(sd: SD, bg_inputs, src_obj_inputs, tsf_obj_inputs, src_hand_inputs, tsf_hand_inputs, T,
                      src_obj_conds=None, src_hand_conds=None, tsf_obj_conds=None, tsf_hand_conds=None,
                      src_armask=None, tsf_armask=None, *, repeat_num: int = 6, n_down: int = 3,
                      spade_layers: Sequence[int] = (1, 1, 0, 0), attn_layers: Sequence[int] = tuple(range(1, 10)),
                      taps: Optional[dict] = None):
    """Generator.forward (generator.py:347-376) -> the reference's 10-tuple.

    ``taps`` (optional dict) receives intermediate tensors for layer-wise
    debugging of the CUDA path.
    """
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}  # base_model.py:112
    if src_obj_conds is None or src_hand_conds is None:
        src_bg = torch.cat([bg_inputs, src_obj_inputs[:, 3:]], 1)
    else:
        src_bg = torch.cat([bg_inputs, src_hand_conds], 1)
    if tsf_obj_conds is None or tsf_hand_conds is None:
        tsf_bg = torch.cat([bg_inputs, tsf_hand_inputs[:, 3:]], 1)
    else:
        tsf_bg = torch.cat([bg_inputs, tsf_hand_conds], 1)
    if src_armask is not None:
        src_bg = torch.cat([src_bg, src_armask], 1)
    if tsf_armask is not None:
        tsf_bg = torch.cat([tsf_bg, tsf_armask], 1)
    src_img_bg = resnet_generator(sd, "bg_model.", src_bg, n_down, repeat_num)
    tsf_img_bg = resnet_generator(sd, "bg_model.", tsf_bg, n_down, repeat_num)

    src_m = _UNet(sd, "src_model.", n_down, repeat_num, spade_layers)
    tsf_m = _UNet(sd, "tsf_model.", n_down, repeat_num, spade_layers)
    obj_m = _UNet(sd, "obj_model.", n_down, repeat_num, spade_layers)

    def warp(layer, s, t):
        if layer in attn_layers:
            return transform(sd, s, T, y=t, attn_prefix=f"attn_{layer}.")
        return transform(sd, s, T)

    # infer_front, generator.py:379-464
    sx, tx = src_m.stem(src_hand_inputs), tsf_m.stem(tsf_hand_inputs)
    s_outs: List[torch.Tensor] = [sx]
    t_outs: List[torch.Tensor] = [tx]
    for i in range(1, n_down + 1):
        sx = src_m.encoder(i, sx, src_hand_conds)
        tx = tsf_m.encoder(i, tx, tsf_hand_conds)
        w = warp(i, sx, tx)
        if taps is not None:
            taps[f"enc{i}.src"], taps[f"enc{i}.tsf_pre"], taps[f"enc{i}.warp"] = sx, tx, w
        tx = tx + w
        s_outs.append(sx)
        t_outs.append(tx)
    for i in range(repeat_num):
        sx = src_m.resnet(i, sx, src_hand_conds)
        tx = tsf_m.resnet(i, tx, tsf_hand_conds)
        w = warp(i + n_down + 1, sx, tx)
        if taps is not None:
            taps[f"res{i}.src"], taps[f"res{i}.tsf_pre"], taps[f"res{i}.warp"] = sx, tx, w
        tx = tx + w
    sy = obj_m.forward(src_obj_inputs, src_obj_conds)
    ty = obj_m.forward(tsf_obj_inputs, tsf_obj_conds)
    sx = src_m.decode(sx, s_outs, src_hand_conds)
    tx = tsf_m.decode(tx, t_outs, tsf_hand_conds)
    if taps is not None:
        taps["src_dec"], taps["tsf_dec"], taps["src_objdec"], taps["tsf_objdec"] = sx, tx, sy, ty
    src_hand = src_m.head("img_reg", sx, torch.tanh)
    src_mask_hand = src_m.head("attetion_reg_hand", sx, torch.sigmoid)
    src_mask_bg = src_m.head("attetion_reg_bg", torch.cat([sx, sy], 1), torch.sigmoid)
    tsf_hand = tsf_m.head("img_reg", tx, torch.tanh)
    tsf_mask_hand = tsf_m.head("attetion_reg_hand", tx, torch.sigmoid)
    tsf_mask_bg = tsf_m.head("attetion_reg_bg", torch.cat([tx, ty], 1), torch.sigmoid)
    src_obj = obj_m.head("img_reg", sy, torch.tanh)
    tsf_obj = obj_m.head("img_reg", ty, torch.tanh)
    return (src_img_bg, tsf_img_bg, src_obj, src_hand, src_mask_bg, src_mask_hand,
            tsf_obj, tsf_hand, tsf_mask_bg, tsf_mask_hand)


def composite(img_bg, obj, hand, mask_bg, mask_hand):
    """models/trainer.py:400-401."""
    return mask_bg * img_bg + (1 - mask_bg) * (obj * mask_hand + hand * (1 - mask_hand))


# --------------------------------------------------- state_dict layout (no weights)
def state_dict_spec(bg_dim=8, img_dim=3, obj_dim=3, img_cond_dim=3, obj_cond_dim=12, conv_dim=64,
                    repeat_num=6, n_down=3, spade_layers=(1, 1, 0, 0), attn_layers=tuple(range(1, 10))):
    """Ordered ``[(key, shape)]`` of the reference Generator's state_dict
    (generator.py:93-345), derived from the constructor arguments only."""
    out = []

    def conv(p, co, ci, k, bias):
        out.append((p + "weight", (co, ci, k, k)))
        if bias:
            out.append((p + "bias", (co,)))

    def inorm(p, c):
        out.append((p + "weight", (c,)))
        out.append((p + "bias", (c,)))

    def spade_p(p, c, s):
        conv(p + "mlp_shared.0.", 128, s, 3, True)
        conv(p + "mlp_gamma.", c, 128, 3, True)
        conv(p + "mlp_beta.", c, 128, 3, True)

    def resblock(p, c):
        conv(p + "main.0.", c, c, 3, False); inorm(p + "main.1.", c)
        conv(p + "main.3.", c, c, 3, False); inorm(p + "main.4.", c)

    # bg_model
    p, i, c = "bg_model.model.", 0, conv_dim
    conv(f"{p}{i}.", c, bg_dim, 7, False); inorm(f"{p}{i+1}.", c); i += 3
    for _ in range(n_down):
        conv(f"{p}{i}.", 2 * c, c, 3, False); inorm(f"{p}{i+1}.", 2 * c); i += 3; c *= 2
    for _ in range(repeat_num):
        resblock(f"{p}{i}.", c); i += 1
    for _ in range(n_down):
        out.append((f"{p}{i}.weight", (c, c // 2, 3, 3))); inorm(f"{p}{i+1}.", c // 2); i += 3; c //= 2
    conv(f"{p}{i}.", 3, c, 7, False)

    def unet(p, c_dim, s_dim, on_obj):
        c = conv_dim
        conv(p + "encoders.0.0.", c, c_dim, 7, False); inorm(p + "encoders.0.1.", c)
        for i in range(1, n_down + 1):
            if spade_layers[0]:
                conv(f"{p}encoders.{i}.conv.", 2 * c, c, 3, False); spade_p(f"{p}encoders.{i}.norm.", 2 * c, s_dim)
            else:
                conv(f"{p}encoders.{i}.0.", 2 * c, c, 3, False); inorm(f"{p}encoders.{i}.1.", 2 * c)
            c *= 2
        for i in range(repeat_num):
            if spade_layers[1] if i < repeat_num // 2 else spade_layers[2]:
                q = f"{p}resnets.{i}."
                conv(q + "conv_0.", c, c, 3, True); conv(q + "conv_1.", c, c, 3, True)
                spade_p(q + "norm_0.", c, s_dim); spade_p(q + "norm_1.", c, s_dim)
            else:
                resblock(f"{p}resnets.{i}.", c)
        dec, skip = [], []
        for i in range(n_down):
            if spade_layers[3]:
                dec.append((f"{p}decoders.{i}.conv.weight", (c, c // 2, 3, 3)))
                n0 = len(out); spade_p(f"{p}decoders.{i}.norm.", c // 2, s_dim); dec.extend(out[n0:]); del out[n0:]
            else:
                dec.append((f"{p}decoders.{i}.0.weight", (c, c // 2, 3, 3)))
                dec.append((f"{p}decoders.{i}.1.weight", (c // 2,))); dec.append((f"{p}decoders.{i}.1.bias", (c // 2,)))
            skip.append((f"{p}skippers.{i}.0.weight", (c // 2, c, 3, 3)))
            skip.append((f"{p}skippers.{i}.1.weight", (c // 2,))); skip.append((f"{p}skippers.{i}.1.bias", (c // 2,)))
            c //= 2
        out.extend(dec); out.extend(skip)
        conv(p + "img_reg.0.", 3, c, 7, False)
        if not on_obj:
            conv(p + "attetion_reg_hand.0.", 1, c, 7, False)
            conv(p + "attetion_reg_bg.0.", 1, 2 * c, 7, False)

    unet("obj_model.", obj_dim, obj_cond_dim, True)
    unet("src_model.", img_dim, img_cond_dim, False)
    unet("tsf_model.", img_dim, img_cond_dim, False)
    chan = {0: conv_dim}
    for i in range(n_down):
        chan[i + 1] = conv_dim * 2 ** (i + 1)
    for i in range(repeat_num):
        chan[i + 1 + n_down] = conv_dim * 2 ** n_down
    for L in attn_layers:
        conv(f"attn_{L}.fully_connect_layer.0.", 128, 2 * chan[L], 5, True)
        conv(f"attn_{L}.fully_connect_layer.2.", 25, 128, 1, True)
    return out


def init_state_dict(seed: int = 0, jitter: float = 0.0, **cfg) -> SD:
    """Random-init weights in the reference layout: conv/convT weights N(0,0.02),
    conv biases 0, InstanceNorm affine (1,0)  (base_network.py:14-25, quirk Q6).
    ``jitter`` > 0 perturbs biases / affine params (N(0,jitter)) so that tests
    exercise them.  Deterministic, but NOT the RNG stream of ``init_weights()``."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in state_dict_spec(**cfg):
        if len(shp) == 4:
            sd[k] = torch.randn(shp, generator=g) * 0.02
        elif k.endswith("bias"):
            sd[k] = torch.randn(shp, generator=g) * jitter
        else:  # 1-D weight: InstanceNorm scale
            sd[k] = 1.0 + torch.randn(shp, generator=g) * jitter
    return sd
