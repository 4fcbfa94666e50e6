"""GeneratorB200's host-side schedule + weight packing, checked on CPU against the oracle
through the torch emulation of the op contracts (tests/emu_ops.py)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from hoig_b200 import synth
from hoig_b200.generator import GeneratorB200, create, parameter_layout
from oracle import generator_ref as gr

from . import emu_ops

SMALL = dict(bg_dim=8, img_dim=3, obj_dim=3, img_cond_dim=3, obj_cond_dim=12, conv_dim=16, repeat_num=6)


def test_state_dict_layout_is_the_reference_layout(golden_dir):
    g = create("generator_spade_attn", **dict(SMALL, conv_dim=64))
    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    got = [[k, list(v.shape)] for k, v in g.state_dict().items()]
    assert got == keys
    assert sum(p.numel() for p in g.parameters()) == 183501729
    assert len(list(g.buffers())) == 0
    assert g.name == "generator"
    assert [n for n, _ in g.named_children()] == ["bg_model", "obj_model", "src_model", "tsf_model"] + [f"attn_{i}" for i in range(1, 10)]


@pytest.mark.parametrize("variant", ["generator_spade_attn", "generator_spade", "generator_base", "generator_spade_attn_tiny"])
def test_layout_matches_oracle_spec_for_all_variants(variant):
    table = {"generator_base": dict(spade_layers=(0, 0, 0, 0), attn_layers=()),
             "generator_spade": dict(spade_layers=(1, 1, 0, 0), attn_layers=()),
             "generator_spade_attn": dict(spade_layers=(1, 1, 0, 0), attn_layers=tuple(range(1, 10))),
             "generator_spade_attn_tiny": dict(spade_layers=(0, 0, 1, 1), attn_layers=tuple(range(1, 10)))}[variant]
    g = create(variant, **SMALL)
    spec = gr.state_dict_spec(**SMALL, **table)
    assert [(k, tuple(v.shape)) for k, v in g.state_dict().items()] == [(k, tuple(s)) for k, s in spec]


def test_init_weights_matches_reference_rule():
    g = create("generator_spade_attn", **SMALL)
    sd0 = {k: v.clone() for k, v in g.state_dict().items()}
    with torch.no_grad():
        for p in g.parameters():
            p.add_(1.0)
    g.init_weights()
    for (name, shape, kind) in g._layout:
        v = g.get_parameter(name)
        if kind == "conv":
            assert abs(v.std().item() - 0.02) < 0.01 and abs(v.mean().item()) < 0.01
        elif kind == "conv_bias":
            assert (v == 0).all()
        else:  # InstanceNorm affine untouched by init_weights (quirk Q6)
            assert torch.equal(v, sd0[name] + 1.0)


@pytest.mark.parametrize("variant,dtype,tol", [("generator_spade_attn", torch.float32, 2e-4),
                                               ("generator_spade", torch.float32, 2e-4),
                                               ("generator_base", torch.float32, 2e-4),
                                               ("generator_spade_attn_tiny", torch.float32, 2e-4)])
def test_schedule_matches_oracle(monkeypatch, variant, dtype, tol):
    emu_ops.install(monkeypatch)
    table = {"generator_base": dict(spade_layers=(0, 0, 0, 0), attn_layers=()),
             "generator_spade": dict(spade_layers=(1, 1, 0, 0), attn_layers=()),
             "generator_spade_attn": dict(spade_layers=(1, 1, 0, 0), attn_layers=tuple(range(1, 10))),
             "generator_spade_attn_tiny": dict(spade_layers=(0, 0, 1, 1), attn_layers=tuple(range(1, 10)))}[variant]
    sd = gr.init_state_dict(seed=0, jitter=0.05, **SMALL, **table)
    g = create(variant, dtype=dtype, **SMALL)
    g.load_state_dict(sd, strict=True)
    inp = synth.generator_inputs(2, seed=1, size=64)
    with torch.no_grad():
        outs = g._forward_impl(**inp)      # the schedule itself; forward() refuses CPU tensors (no CPU path in the product)
    with torch.no_grad():
        ref = gr.generator_forward(sd, **inp, **table)
    assert len(outs) == 10
    for i, (a, b) in enumerate(zip(outs, ref)):
        assert a.shape == b.shape and a.dtype == torch.float32
        assert (a - b).abs().max().item() <= tol, (i, (a - b).abs().max().item())


def test_schedule_bf16_emulation_within_rel_l2(monkeypatch):
    emu_ops.install(monkeypatch)
    table = dict(spade_layers=(1, 1, 0, 0), attn_layers=tuple(range(1, 10)))
    sd = gr.init_state_dict(seed=0, jitter=0.05, **SMALL, **table)
    g = create("generator_spade_attn", dtype=torch.bfloat16, **SMALL)
    g.load_state_dict(sd)
    inp = synth.generator_inputs(1, seed=1, size=64)
    with torch.no_grad():
        outs = g._forward_impl(**inp)      # the schedule itself; forward() refuses CPU tensors (no CPU path in the product)
    with torch.no_grad():
        ref = gr.generator_forward(sd, **inp, **table)
    for i, (a, b) in enumerate(zip(outs, ref)):
        rel = ((a - b).norm() / b.norm()).item()
        assert rel <= 3e-2, (i, rel)   # bf16 storage at 64x64 / conv_dim 16; the GPU gate (1e-2) is on the full config


def test_schedule_f16_emulation_within_rel_l2(monkeypatch):
    emu_ops.install(monkeypatch)
    table = dict(spade_layers=(1, 1, 0, 0), attn_layers=tuple(range(1, 10)))
    sd = gr.init_state_dict(seed=0, jitter=0.05, **SMALL, **table)
    g = create("generator_spade_attn", dtype=torch.float16, **SMALL)
    g.load_state_dict(sd)
    inp = synth.generator_inputs(1, seed=1, size=64)
    with torch.no_grad():
        outs = g._forward_impl(**inp)      # the schedule itself; forward() refuses CPU tensors (no CPU path in the product)
    with torch.no_grad():
        ref = gr.generator_forward(sd, **inp, **table)
    for i, (a, b) in enumerate(zip(outs, ref)):
        assert ((a - b).norm() / b.norm()).item() <= 1e-2, i


def test_packed_cache_invalidates_on_update(monkeypatch):
    emu_ops.install(monkeypatch)
    g = create("generator_base", dtype=torch.float32, **SMALL)
    w1 = g._w("bg_model.model.0.weight")
    assert g._w("bg_model.model.0.weight") is w1
    with torch.no_grad():
        g.get_parameter("bg_model.model.0.weight").mul_(2.0)
    w2 = g._w("bg_model.model.0.weight")
    assert w2 is not w1 and torch.allclose(w2, w1 * 2)


def test_commuted_attention_equals_block_extractor_formulation():
    """The tensor-core attention (two dense 5x5 convs over replicate-padded rasters + bilinear interpolation of the
    source conv) is the same sum as extract_attn.py:24-28 over BlockExtractor taps, including taps clamped at the
    image border and flows that leave the image."""
    from hoig_b200.packing import pack_conv_weight
    g = torch.Generator().manual_seed(5)
    n, h, c, k, hid = 2, 12, 64, 5, 128
    src, tgt = torch.randn(n, h, h, c, generator=g), torch.randn(n, h, h, c, generator=g)
    flow = torch.randn(n, h, h, 2, generator=g) * 4.0
    flow[0, 0, 0] = torch.tensor([-30.0, 25.0]); flow[1, 5, 5] = torch.tensor([40.0, -0.5]); flow[1, 3, 2] = torch.tensor([2.0, -3.0])
    w0 = torch.randn(hid, 2 * c, k, k, generator=g) * 0.02
    b1, w2, b2 = torch.randn(hid, generator=g) * 0.1, torch.randn(k * k, hid, generator=g) * 0.2, torch.randn(k * k, generator=g) * 0.1
    # reference association: taps -> k5s5 conv (as a GEMM over the unfolded taps) -> finish
    unf = emu_ops.attn_unfold(src, tgt, flow, torch.empty(n, h, h, 2 * k * k * c), k)
    wfull = w0.permute(0, 2, 3, 1).reshape(hid, k * k * 2 * c)          # K order (tap, [tgt C | src C])
    hidden = F.leaky_relu(unf @ wfull.t() + b1, 0.01)
    ref = emu_ops.attn_finish(hidden, w2, b2, src, flow, tgt, torch.empty(n, h, h, c), k, unfold=unf)
    # commuted association
    r = k // 2
    tpad = emu_ops.replicate_pad(tgt, torch.empty(n, h + 2 * r, h + 2 * r, c), r)
    spad = emu_ops.replicate_pad(src, torch.empty(n, h + 4 * r, h + 4 * r, c), 2 * r)
    wt, ws = pack_conv_weight(w0[:, :c], torch.float32), pack_conv_weight(w0[:, c:], torch.float32)
    gt, gs = emu_ops.conv2d_halo([(tpad, wt, torch.empty(n, h + 2 * r, h + 2 * r, hid)),
                                  (spad, ws, torch.empty(n, h + 4 * r, h + 4 * r, hid))], k, k, hid)
    out = emu_ops.attn_combine(gt, gs, b1, w2, b2, src, flow, tgt, torch.empty(n, h, h, c), k)
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= 2e-5


def test_forward_refuses_cpu_tensors():
    g = create("generator_spade_attn", **SMALL)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g(**synth.generator_inputs(1, seed=1, size=64))


def test_weight_cache_invalidation_rules():
    """replacing a Parameter object, load_state_dict (also assign=True) and .to() must drop the derived copies;
    in-place updates are seen through the version counter."""
    g = create("generator_spade_attn", **SMALL)
    name = "src_model.img_reg.0.weight"
    p0 = g._p(name)
    sig0 = g._weights_signature()
    g._pcache["probe"] = ((), 1)
    with torch.no_grad():
        p0.mul_(2.0)
    assert g._weights_signature() != sig0
    g.src_model.img_reg._modules["0"].weight = torch.nn.Parameter(torch.zeros_like(p0))
    assert g._p(name) is not p0 and "probe" not in g._pcache
    g._pcache["probe"] = ((), 1)
    g.load_state_dict({k: v.clone() for k, v in g.state_dict().items()}, assign=True)
    assert "probe" not in g._pcache and g._p(name) is g.get_parameter(name)
    g._pcache["probe"] = ((), 1)
    g.train(); g.eval()
    assert "probe" not in g._pcache
    g._pcache["probe"] = ((), 1)
    g.double()
    assert "probe" not in g._pcache
