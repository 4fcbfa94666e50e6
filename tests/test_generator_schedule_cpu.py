"""GeneratorB200's host-side schedule + weight packing, checked on CPU against the oracle
through the torch emulation of the op contracts (tests/emu_ops.py)."""
import json
import os

import numpy as np
import pytest
import torch

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
    outs = g(**inp)
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
    outs = g(**inp)
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
    outs = g(**inp)
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
