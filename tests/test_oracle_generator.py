"""Pins oracle/generator_ref.py against fixtures produced by the unmodified
reference Generator (tests/golden/make_golden.py) and its state_dict layout."""
import json
import os

import numpy as np
import pytest
import torch

from hoig_b200 import synth
from oracle import generator_ref as gr

SMALL = dict(bg_dim=8, img_dim=3, obj_dim=3, img_cond_dim=3, obj_cond_dim=12, conv_dim=16, repeat_num=6)
FULL = dict(SMALL, conv_dim=64)


def _check(golden_dir, name, cfg, size, B, stride, tol):
    g = np.load(os.path.join(golden_dir, f"generator_{name}.npz"))
    sd = gr.init_state_dict(seed=0, jitter=0.05, **cfg)
    inp = synth.generator_inputs(B, seed=1, size=size)
    with torch.no_grad():
        outs = gr.generator_forward(sd, **inp)
    assert len(outs) == 10
    for i, o in enumerate(outs):
        ref = g[f"out{i}_sample"]
        got = o[:, :, ::stride, ::stride].numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= tol, (i, np.abs(got - ref).max())
        s = g[f"out{i}_sum"]
        assert abs(o.double().sum().item() - s[0]) <= tol * o.numel() * 0.05 + 1e-3
        assert abs(o.double().abs().sum().item() - s[1]) <= tol * o.numel() * 0.05 + 1e-3


def test_restatement_matches_reference_small(golden_dir):
    # the restatement executes the same torch ops in the same order: tight tolerance
    _check(golden_dir, "small", SMALL, 64, 2, 4, 2e-5)


@pytest.mark.slow
def test_restatement_matches_reference_full(golden_dir):
    _check(golden_dir, "full", FULL, 256, 1, 8, 5e-5)


def test_state_dict_layout_matches_reference(golden_dir):
    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    spec = gr.state_dict_spec(**FULL)
    assert len(spec) == 425
    assert [[k, list(s)] for k, s in spec] == keys
    assert sum(int(np.prod(s)) for _, s in spec) == 183501729


def test_block_extract_torch_vs_c_oracle():
    import oracle
    g = torch.Generator().manual_seed(3)
    src = torch.randn(2, 5, 9, 7, generator=g)
    flow = torch.randn(2, 2, 9, 7, generator=g) * 3
    a = gr.block_extract(src, flow, 5).numpy()
    b = oracle.block_extract(src.numpy(), flow.numpy(), 5)
    assert np.abs(a - b).max() < 1e-6
    # zero flow == replicate-padded unfold
    z = oracle.block_extract(src.numpy(), np.zeros_like(flow.numpy()), 5)
    u = torch.nn.functional.unfold(torch.nn.functional.pad(src, (2, 2, 2, 2), mode="replicate"), 5)
    u = u.reshape(2, 5, 5, 5, 9, 7).permute(0, 1, 4, 2, 5, 3).reshape(2, 5, 45, 35)
    assert np.array_equal(z, u.numpy())


def test_local_attn_reshape_kat():
    """thirdparty/local_attn_reshape/test_local_attn_reshape.py:30-44 prints
    out[0,0,:3,:3] for the arange(9) pattern; the implied answer is 0..8."""
    import oracle
    x = np.arange(9, dtype=np.float32).reshape(1, 9, 1, 1).repeat(14, 2).repeat(10, 3)
    out = oracle.local_attn_reshape(x, 3)
    assert out.shape == (1, 1, 42, 30)
    assert np.array_equal(out[0, 0, :3, :3], np.arange(9, dtype=np.float32).reshape(3, 3))
    assert np.array_equal(out, gr.local_attn_reshape(torch.from_numpy(x), 3).numpy())


def test_oracle_block_extract_backward_matches_autograd():
    """The C restatement of kernel_block_extractor_backward (block_extractor_kernel.cu:86-166) against torch autograd through the
    torch restatement of the forward (whose bilinear weights are differentiable in the flow exactly the way the kernel's analytic
    formula is), and the LocalAttnReshape backward against the autograd of pixel_shuffle."""
    import oracle
    g = torch.Generator().manual_seed(0)
    B, C, H, W, k = 2, 3, 7, 6, 5
    src = torch.randn(B, C, H, W, generator=g, dtype=torch.float64, requires_grad=True)
    flow = (torch.randn(B, 2, H, W, generator=g, dtype=torch.float64) * 2.5).requires_grad_(True)
    gout = torch.randn(B, C, k * H, k * W, generator=g, dtype=torch.float64)
    out = gr.block_extract(src, flow, k)
    out.backward(gout)
    gs, gf = oracle.block_extract_backward(src.detach().float().numpy(), flow.detach().float().numpy(), gout.float().numpy(), k)
    assert np.abs(gs - src.grad.numpy()).max() <= 2e-4 * max(1.0, np.abs(src.grad.numpy()).max())
    assert np.abs(gf - flow.grad.numpy()).max() <= 2e-4 * max(1.0, np.abs(flow.grad.numpy()).max())
    x = torch.randn(2, k * k, 4, 3, generator=g, requires_grad=True)
    go = torch.randn(2, 1, 4 * k, 3 * k, generator=g)
    gr.local_attn_reshape(x, k).backward(go)
    assert np.array_equal(oracle.local_attn_reshape_backward(go.numpy(), k), x.grad.numpy())
