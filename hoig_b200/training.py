"""Training path: differentiable generator forward, PatchGAN discriminator, and the
generator + discriminator step of ``Trainer.optimize_parameters`` (models/trainer.py:417-481) with the DDP gradient all-reduce of
``train_ddp.py`` / ``models/trainer.py:237-252``.

Everything heavy runs on this repo's kernels through ``hoig_b200.autograd`` (fp32: implicit-GEMM conv forward / data gradient,
``hoig_conv2d_wgrad_f32``, InstanceNorm forward / backward, BlockExtractor and LocalAttnReshape forward / backward); torch supplies
autograd bookkeeping, elementwise glue, the Adam optimiser and ``torch.distributed``.  The op order is the reference's
(generator.py:347-491, spade.py:24-38, extract_attn.py:23-29), so the parameters, their gradients and the checkpoint layout are
the reference's.  The fused 16-bit inference schedule (``GeneratorB200._forward_impl``) is not differentiable and is not used here.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import autograd as ag
from . import ops

ATTN_K = 5


# ------------------------------------------------------------------------------------------------ generator (row N3)
def _inorm(g, x, p):
    return ag.instance_norm(x, g.get_parameter(p + "weight"), g.get_parameter(p + "bias"))


def _spade(g, p, x, seg):
    """spade.py:24-38: instance norm (no affine), nearest-resized segmentation map -> shared 3x3 conv + ReLU -> gamma / beta."""
    normalized = ag.instance_norm(x)
    seg = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    actv = torch.relu(ag.conv2d(seg, g.get_parameter(p + "mlp_shared.0.weight"), g.get_parameter(p + "mlp_shared.0.bias"), 1, 1))
    gamma = ag.conv2d(actv, g.get_parameter(p + "mlp_gamma.weight"), g.get_parameter(p + "mlp_gamma.bias"), 1, 1)
    beta = ag.conv2d(actv, g.get_parameter(p + "mlp_beta.weight"), g.get_parameter(p + "mlp_beta.bias"), 1, 1)
    return normalized * (1 + gamma) + beta


def _conv_in_relu(g, p, x, stride=1, padding=1, transposed=False):
    if transposed:
        h = ag.conv_transpose2d(x, g.get_parameter(p + "0.weight"))
    else:
        h = ag.conv2d(x, g.get_parameter(p + "0.weight"), None, stride, padding)
    return torch.relu(_inorm(g, h, p + "1."))


def _residual_block(g, p, x):
    """generator.py:9-32."""
    h = torch.relu(_inorm(g, ag.conv2d(x, g.get_parameter(p + "main.0.weight"), None, 1, 1), p + "main.1."))
    h = _inorm(g, ag.conv2d(h, g.get_parameter(p + "main.3.weight"), None, 1, 1), p + "main.4.")
    return x + h


def _spade_residual_block(g, p, x, seg):
    """generator.py:35-71."""
    dx = ag.conv2d(torch.relu(_spade(g, p + "norm_0.", x, seg)), g.get_parameter(p + "conv_0.weight"), g.get_parameter(p + "conv_0.bias"), 1, 1)
    dx = ag.conv2d(torch.relu(_spade(g, p + "norm_1.", dx, seg)), g.get_parameter(p + "conv_1.weight"), g.get_parameter(p + "conv_1.bias"), 1, 1)
    return x + dx


def _bg(g, x):
    """ResNetGenerator.forward, generator.py:93-135."""
    p, i = "bg_model.model.", 0
    h = torch.relu(_inorm(g, ag.conv2d(x, g.get_parameter(f"{p}{i}.weight"), None, 1, 3), f"{p}{i + 1}.")); i += 3
    for _ in range(g.n_down):
        h = torch.relu(_inorm(g, ag.conv2d(h, g.get_parameter(f"{p}{i}.weight"), None, 2, 1), f"{p}{i + 1}.")); i += 3
    for _ in range(g.repeat_num):
        h = _residual_block(g, f"{p}{i}.", h); i += 1
    for _ in range(g.n_down):
        h = torch.relu(_inorm(g, ag.conv_transpose2d(h, g.get_parameter(f"{p}{i}.weight")), f"{p}{i + 1}.")); i += 3
    return torch.tanh(ag.conv2d(h, g.get_parameter(f"{p}{i}.weight"), None, 1, 3))


class _UNet:
    """ResUnetGenerator (generator.py:138-315) over a parameter-name prefix."""

    def __init__(self, g, p):
        self.g, self.p = g, p

    def stem(self, x):
        return _conv_in_relu(self.g, self.p + "encoders.0.", x, padding=3)

    def encoder(self, i, x, seg):
        g, p = self.g, f"{self.p}encoders.{i}."
        if g.spade_layers[0]:
            return torch.relu(_spade(g, p + "norm.", ag.conv2d(x, g.get_parameter(p + "conv.weight"), None, 2, 1), seg))
        return _conv_in_relu(g, p, x, stride=2)

    def resnet(self, i, x, seg):
        p = f"{self.p}resnets.{i}."
        return _spade_residual_block(self.g, p, x, seg) if self.g._is_spade_res(i) else _residual_block(self.g, p, x)

    def decode(self, x, enc_outs, seg):
        g, d = self.g, x
        for i in range(g.n_down):
            p = f"{self.p}decoders.{i}."
            if g.spade_layers[3]:
                d = torch.relu(_spade(g, p + "norm.", ag.conv_transpose2d(d, g.get_parameter(p + "conv.weight")), seg))
            else:
                d = _conv_in_relu(g, p, d, transposed=True)
            d = torch.cat([enc_outs[g.n_down - 1 - i], d], 1)
            d = _conv_in_relu(g, f"{self.p}skippers.{i}.", d)
        return d

    def forward(self, x, seg):
        h = self.stem(x)
        outs = [h]
        for i in range(1, self.g.n_down + 1):
            h = self.encoder(i, h, seg)
            outs.append(h)
        for i in range(self.g.repeat_num):
            h = self.resnet(i, h, seg)
        return self.decode(h, outs, seg)

    def head(self, name, x, act):
        return act(ag.conv2d(x, self.g.get_parameter(f"{self.p}{name}.0.weight"), None, 1, 3))


def _attn_warp(g, p, src, tgt, flow):
    """extract_attn.py:23-29 with the reference's own op order (the 25x block tensors ARE materialised here, as in the reference)."""
    k = ATTN_K
    bs = ag.block_extract(src, flow, k)
    bt = ag.block_extract(tgt, torch.zeros_like(flow), k)
    h = ag.conv2d(torch.cat((bt, bs), 1), g.get_parameter(p + "fully_connect_layer.0.weight"),
                  g.get_parameter(p + "fully_connect_layer.0.bias"), k, 0)
    h = F.leaky_relu(h, 0.01)
    a = F.softmax(ag.conv2d(h, g.get_parameter(p + "fully_connect_layer.2.weight"), g.get_parameter(p + "fully_connect_layer.2.bias"), 1, 0), 1)
    a = ag.local_attn_reshape(a, k)
    return F.avg_pool2d(a * bs, k, k)


def generator_forward_train(g, bg_inputs, src_obj_inputs, tsf_obj_inputs, src_hand_inputs, tsf_hand_inputs, T,
                            src_obj_conds=None, src_hand_conds=None, tsf_obj_conds=None, tsf_hand_conds=None,
                            src_armask=None, tsf_armask=None):
    """Differentiable ``Generator.forward`` (generator.py:347-376) for ``GeneratorB200`` ``g``: the reference's 10-tuple."""
    def cat(parts):
        return torch.cat([t for t in parts if t is not None], 1)

    src_bg = cat([bg_inputs, src_obj_inputs[:, 3:] if (src_obj_conds is None or src_hand_conds is None) else src_hand_conds, src_armask])
    tsf_bg = cat([bg_inputs, tsf_hand_inputs[:, 3:] if (tsf_obj_conds is None or tsf_hand_conds is None) else tsf_hand_conds, tsf_armask])
    src_img_bg, tsf_img_bg = _bg(g, src_bg), _bg(g, tsf_bg)
    src_m, tsf_m, obj_m = _UNet(g, "src_model."), _UNet(g, "tsf_model."), _UNet(g, "obj_model.")
    nd = g.n_down
    flows: Dict[int, torch.Tensor] = {}
    Tf = T.detach().float().contiguous()

    def warp(layer, s, t):
        h = s.shape[2]
        attn = layer in g.attn_layers
        if (h, attn) not in flows:                       # resize_trans (+ identity subtraction for the attention flow), no gradient
            with torch.no_grad():
                flows[(h, attn)] = ops.resize_flow(Tf, h, subtract_identity=attn)
        fl = flows[(h, attn)]
        if attn:
            return _attn_warp(g, f"attn_{layer}.", s, t, fl.permute(0, 3, 1, 2).contiguous())
        return F.grid_sample(s, fl, mode="bilinear", padding_mode="zeros", align_corners=False)   # generator.py:475-478 (torch glue)

    sx, tx = src_m.stem(src_hand_inputs), tsf_m.stem(tsf_hand_inputs)
    s_outs, t_outs = [sx], [tx]
    for i in range(1, nd + 1):
        sx = src_m.encoder(i, sx, src_hand_conds)
        tx = tsf_m.encoder(i, tx, tsf_hand_conds)
        tx = tx + warp(i, sx, tx)
        s_outs.append(sx); t_outs.append(tx)
    for i in range(g.repeat_num):
        sx = src_m.resnet(i, sx, src_hand_conds)
        tx = tsf_m.resnet(i, tx, tsf_hand_conds)
        tx = tx + warp(i + nd + 1, sx, tx)
    sy = obj_m.forward(src_obj_inputs, src_obj_conds)
    ty = obj_m.forward(tsf_obj_inputs, tsf_obj_conds)
    sx = src_m.decode(sx, s_outs, src_hand_conds)
    tx = tsf_m.decode(tx, t_outs, tsf_hand_conds)
    src_hand = src_m.head("img_reg", sx, torch.tanh)
    src_mask_hand = src_m.head("attetion_reg_hand", sx, torch.sigmoid)
    src_mask_bg = src_m.head("attetion_reg_bg", torch.cat([sx, sy], 1), torch.sigmoid)
    tsf_hand = tsf_m.head("img_reg", tx, torch.tanh)
    tsf_mask_hand = tsf_m.head("attetion_reg_hand", tx, torch.sigmoid)
    tsf_mask_bg = tsf_m.head("attetion_reg_bg", torch.cat([tx, ty], 1), torch.sigmoid)
    src_obj = obj_m.head("img_reg", sy, torch.tanh)
    tsf_obj = obj_m.head("img_reg", ty, torch.tanh)
    return (src_img_bg, tsf_img_bg, src_obj, src_hand, src_mask_bg, src_mask_hand, tsf_obj, tsf_hand, tsf_mask_bg, tsf_mask_hand)


# ------------------------------------------------------------------------------------------------ discriminator (row N4)
class PatchDiscriminatorB200(nn.Module):
    """Drop-in for ``PatchDiscriminator`` (models/networks/discriminator.py:8-57): same constructor, ``forward`` and state_dict
    keys (``model.{0,2,5,...}.weight/bias``), 4x4 convs + LeakyReLU(0.2) + InstanceNorm2d(affine=False) on this repo's kernels."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_type="instance", use_sigmoid=False):
        super().__init__()
        if norm_type != "instance":
            raise NotImplementedError("PatchDiscriminatorB200: the shipped configuration uses norm_type='instance' (options/base_options.py:48)")
        self._name = "discriminator_patch_gan"
        self.use_sigmoid = use_sigmoid
        self.model = nn.Module()
        self.layers: List[tuple] = []      # (sequential index, stride, has_norm, has_act)
        idx, nf, nf_prev = 0, 1, 1

        def add(i, cin, cout):
            m = nn.Module()
            m.register_parameter("weight", nn.Parameter(torch.empty(cout, cin, 4, 4)))
            m.register_parameter("bias", nn.Parameter(torch.zeros(cout)))     # use_bias is True for InstanceNorm (discriminator.py:23)
            self.model.add_module(str(i), m)

        add(idx, input_nc, ndf); self.layers.append((idx, 2, False, True)); idx += 2
        for n in range(1, n_layers):
            nf_prev, nf = nf, min(2 ** n, 8)
            add(idx, ndf * nf_prev, ndf * nf); self.layers.append((idx, 2, True, True)); idx += 3
        nf_prev, nf = nf, min(2 ** n_layers, 8)
        add(idx, ndf * nf_prev, ndf * nf); self.layers.append((idx, 1, True, True)); idx += 3
        add(idx, ndf * nf, 1); self.layers.append((idx, 1, False, False))
        self.init_weights()

    @property
    def name(self):
        return self._name

    def init_weights(self):
        """base_network.py:14-25."""
        with torch.no_grad():
            for m in self.model.children():
                m.weight.normal_(0.0, 0.02)
                m.bias.zero_()

    def forward(self, x):
        for idx, stride, has_norm, has_act in self.layers:
            m = getattr(self.model, str(idx))
            x = ag.conv2d(x, m.weight, m.bias, stride, 1)
            if has_norm:
                x = ag.instance_norm(x)
            if has_act:
                x = F.leaky_relu(x, 0.2)
        return torch.sigmoid(x) if self.use_sigmoid else x


# ------------------------------------------------------------------------------------------------ the training step
def composite(img_bg, obj, hand, mask_bg, mask_hand):
    """models/trainer.py:400-401 (differentiable torch form)."""
    return mask_bg * img_bg + (1 - mask_bg) * (obj * mask_hand + hand * (1 - mask_hand))


def _loss_smooth(mat):
    """trainer.py:470-472."""
    return (mat[:, :, :, :-1] - mat[:, :, :, 1:]).abs().mean() + (mat[:, :, :-1, :] - mat[:, :, 1:, :]).abs().mean()


def allreduce_gradients(params, world_size: Optional[int] = None, bucket_bytes: int = 64 << 20) -> int:
    """DDP's job (models/trainer.py:237-252 wraps G and D in DistributedDataParallel): average the gradients over the ranks with
    bucketed all-reduces (NCCL over NVLink on GPUs, gloo in the CPU tests).  Returns the number of bytes reduced."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world_size = world_size or dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    total, bucket, size = 0, [], 0

    def flush():
        nonlocal bucket, size, total
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat)
        flat.div_(world_size)
        off = 0
        for t in bucket:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()
        total += flat.numel() * flat.element_size()
        bucket, size = [], 0

    for t in grads:
        bucket.append(t)
        size += t.numel() * t.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return total


class TrainStep:
    """One ``Trainer.optimize_parameters`` iteration (models/trainer.py:417-481) for a generator / discriminator pair.

    Losses as the reference: LSGAN terms (``_compute_loss_D``), L1 reconstruction of the source, the target term, mask MSE and mask
    smoothness, weighted by the ``lambda_*`` options.  The reference's target term is a VGG19 perceptual loss on ImageNet weights
    (models/networks/vgg19.py:56), which cannot be loaded offline; ``tsf_loss`` defaults to L1 on the target image and accepts any
    callable ``(fake, real) -> tensor`` (e.g. a VGG loss) instead."""

    def __init__(self, G, D, lr_G=2e-4, lr_D=2e-4, betas=(0.5, 0.999), lambda_rec=10.0, lambda_tsf=10.0, lambda_D_prob=1.0,
                 lambda_mask=1.0, lambda_mask_smooth=1.0, tsf_loss=None):
        self.G, self.D = G, D
        self.opt_G = torch.optim.Adam(G.parameters(), lr=lr_G, betas=betas)      # trainer.py:275-278
        self.opt_D = torch.optim.Adam(D.parameters(), lr=lr_D, betas=betas)
        self.lam = dict(rec=lambda_rec, tsf=lambda_tsf, d=lambda_D_prob, mask=lambda_mask, smooth=lambda_mask_smooth)
        self.tsf_loss = tsf_loss or (lambda fake, real: (fake - real).abs().mean())
        self.allreduce_bytes = 0
        self.allreduce_ms = 0.0          # device time of the gradient all-reduces of the last step (CUDA events)

    def _allreduce(self, params):
        if not torch.cuda.is_available():
            return allreduce_gradients(params)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = allreduce_gradients(params)
        e1.record()
        self._events.append((e0, e1))
        return n

    def __call__(self, gen_kwargs, real_src, real_tsf, bg_mask, hand_mask, train_D: bool = True):
        """gen_kwargs: the Generator.forward keyword arguments (HandRecoveryFlowB200 output); real_* (B,3,H,W); bg_mask / hand_mask
        (2B,1,H,W) = cat of the source and target crop masks (trainer.py:359-360).  Returns the dict of loss values."""
        lam = self.lam
        self._events = []
        outs = self.G(**gen_kwargs)
        (src_bg, tsf_bg, src_obj, src_hand, src_mbg, src_mh, tsf_obj, tsf_hand, tsf_mbg, tsf_mh) = outs
        fake_src = composite(src_bg, src_obj, src_hand, src_mbg, src_mh)
        fake_tsf = composite(tsf_bg, tsf_obj, tsf_hand, tsf_mbg, tsf_mh)
        masks_bg, masks_hand = torch.cat([src_mbg, tsf_mbg], 0), torch.cat([src_mh, tsf_mh], 0)
        cond = [gen_kwargs["tsf_obj_conds"], gen_kwargs["tsf_hand_conds"]]
        if gen_kwargs.get("tsf_armask") is not None:
            cond.append(gen_kwargs["tsf_armask"])
        tsf_cond = torch.cat(cond, 1)
        # ---- generator (trainer.py:436-461)
        d_fake = self.D(torch.cat([fake_tsf, tsf_cond], 1))
        l_adv = ((d_fake - 0) ** 2).mean() * lam["d"]
        l_rec = (fake_src - real_src).abs().mean() * lam["rec"]
        l_tsf = self.tsf_loss(fake_tsf, real_tsf).mean() * lam["tsf"]
        l_mask = (F.mse_loss(masks_bg, bg_mask) + F.mse_loss(masks_hand, hand_mask)) * lam["mask"]
        l_smooth = (_loss_smooth(masks_bg) + _loss_smooth(masks_hand)) * lam["smooth"] if lam["smooth"] else masks_bg.new_zeros(())
        loss_G = l_adv + l_rec + l_tsf + l_mask + l_smooth
        self.opt_G.zero_grad(set_to_none=True)
        self.opt_D.zero_grad(set_to_none=True)
        loss_G.backward()
        self.allreduce_bytes = self._allreduce(list(self.G.parameters()))
        self.opt_G.step()
        losses = dict(g_adv=l_adv.item(), g_rec=l_rec.item(), g_tsf=l_tsf.item(), g_mask=l_mask.item(), g_mask_smooth=float(l_smooth.detach()))
        # ---- discriminator (trainer.py:463-481)
        if train_D:
            self.opt_D.zero_grad(set_to_none=True)
            d_real = self.D(torch.cat([real_tsf, tsf_cond], 1))
            d_fake = self.D(torch.cat([fake_tsf.detach(), tsf_cond], 1))
            loss_D = (((d_real - 1) ** 2).mean() + ((d_fake + 1) ** 2).mean()) * lam["d"]
            loss_D.backward()
            self.allreduce_bytes += self._allreduce(list(self.D.parameters()))
            self.opt_D.step()
            losses.update(d_real=d_real.mean().item(), d_fake=d_fake.mean().item(), loss_D=loss_D.item())
        if self._events:
            torch.cuda.synchronize()
            self.allreduce_ms = sum(a.elapsed_time(b) for a, b in self._events)
        if hasattr(self.G, "refresh_weights"):
            self.G.refresh_weights()           # the packed inference copies are stale after the optimiser step
        return losses
