"""Stage R8 (UV-texture warp, utils/nmr.py:973-1100 + models/trainer.py:83-87): CUDA kernels through the C ABI against the
oracle restatement (oracle/geometry_ref.py) on seeded random meshes / atlases."""
import pytest
import torch
import torch.nn.functional as F

from hoig_b200 import ops, renderer
from oracle import geometry_ref as gref

pytestmark = pytest.mark.gpu


def _scene(B, Fc, size, hu, wu, seed):
    g = torch.Generator().manual_seed(seed)
    src_faces = (torch.rand(B, Fc, 3, 3, generator=g) * 2.4 - 1.2)
    fim_uv = torch.randint(-1, Fc, (hu, wu), generator=g, dtype=torch.int32)
    w = torch.rand(hu, wu, 3, generator=g) + 0.05
    wim_uv = w / w.sum(-1, keepdim=True)
    im = torch.rand(B, 3, size, size, generator=g) * 2 - 1
    # source face-index maps: random, with the true face planted near the warp target of ~half of the atlas pixels
    src_fim = torch.randint(-1, Fc, (B, size, size), generator=g, dtype=torch.int32)
    f2v = src_faces[..., :2].clone()
    f2v[..., 1] *= -1
    _, T, _ = gref.texture_backward_warp(im, f2v, torch.full((B, size, size), -1, dtype=torch.int32), fim_uv, wim_uv)
    t = ((T + 1) / 2.0 * float(size - 1)).long().clamp(0, size - 1)
    plant = (torch.rand(B, hu, wu, generator=g) < 0.5) & (fim_uv[None] != -1)
    for b in range(B):
        ys, xs = t[b, ..., 1][plant[b]], t[b, ..., 0][plant[b]]
        src_fim[b, ys, xs] = fim_uv[plant[b]]
    return src_faces, f2v, fim_uv, wim_uv, src_fim, im


@pytest.mark.parametrize("size,hu,wu,x0", [(64, 32, 80, 48), (256, 256, 640, 384)], ids=["small", "full_atlas"])
def test_texture_backward_warp_vs_oracle(size, hu, wu, x0):
    B, Fc = 2, 300
    src_faces, f2v, fim_uv, wim_uv, src_fim, im = _scene(B, Fc, size, hu, wu, 1)
    g = torch.Generator().manual_seed(5)
    obj_tex = torch.rand(hu, wu - x0, 3, generator=g)
    T, O = ops.uv_backward_warp(src_faces.cuda(), fim_uv.cuda(), wim_uv.cuda(), src_fim.cuda())
    for tex in (None, obj_tex):
        ref, T_ref, O_ref = gref.texture_backward_warp(im, f2v, src_fim, fim_uv, wim_uv, tex, x0)
        out = renderer.texture_backward_warp(im.cuda(), src_faces.cuda(), src_fim.cuda(), fim_uv.cuda(), wim_uv.cuda(),
                                             None if tex is None else tex.cuda(), x0)
        torch.cuda.synchronize()
        assert (out.cpu() - ref).abs().max().item() <= 2e-6
    assert torch.equal(T.cpu(), T_ref)                      # same products and sums in the same order: bit-exact
    # the oracle returns the opened occlusion map; the raw one is recomputed here from T
    raw = torch.zeros(B, hu * wu)
    for b in range(B):
        f = fim_uv.long().reshape(-1)
        ex = f != -1
        t11 = ((T_ref[b].reshape(-1, 2)[ex] + 1) / 2.0 * float(size - 1)).long().clamp(0, size - 1)
        vis = torch.zeros(int(ex.sum()), dtype=torch.bool)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                tt = (t11 + torch.tensor([dx, dy])).clamp(0, size - 1)
                vis |= src_fim[b].long().reshape(-1)[tt[:, 1] * size + tt[:, 0]] == f[ex]
        raw[b, ex] = 1 - vis.float()
    assert torch.equal(O.cpu().reshape(B, -1), raw)
    assert 0.2 < raw[:, (fim_uv.reshape(-1) != -1)].mean().item() < 0.8     # both outcomes are exercised


def test_sample_from_texture_dense_and_render_vs_oracle():
    B, Fc, size, hu, wu = 3, 200, 96, 64, 160
    g = torch.Generator().manual_seed(9)
    uv = torch.rand(Fc, 3, 2, generator=g) * 2.2 - 1.1
    fim = torch.randint(-1, Fc, (B, size, size), generator=g, dtype=torch.int32)
    w = torch.rand(B, size, size, 3, generator=g) + 0.05
    wim = w / w.sum(-1, keepdim=True)
    tex = torch.rand(B, 3, hu, wu, generator=g)
    T = renderer.sample_from_texture_dense(fim.cuda(), wim.cuda(), uv.cuda())
    assert torch.equal(T.cpu(), gref.sample_from_texture_dense(fim, wim, uv))
    out = renderer.render_from_texture(tex.cuda(), fim.cuda(), wim.cuda(), uv.cuda())
    ref = gref.render_from_texture(tex, fim, wim, uv)
    assert (out.cpu() - ref).abs().max().item() <= 2e-6


@pytest.mark.parametrize("align", [False, True])
def test_grid_sample_nchw_vs_torch(align):
    g = torch.Generator().manual_seed(3)
    im = torch.randn(2, 5, 37, 53, generator=g)
    grid = torch.rand(2, 29, 31, 2, generator=g) * 2.6 - 1.3          # includes out-of-range samples (zeros padding)
    grid[0, 0, 0] = torch.tensor([-1.0, 1.0]); grid[0, 0, 1] = torch.tensor([1.0, -1.0]); grid[1, 3, 3] = torch.tensor([-2.0, -2.0])
    out = ops.grid_sample_nchw(im.cuda(), grid.cuda(), align)
    ref = F.grid_sample(im, grid, mode="bilinear", padding_mode="zeros", align_corners=align)
    assert (out.cpu() - ref).abs().max().item() <= 2e-6


def test_texture_ops_reject_bad_arguments():
    with pytest.raises(ValueError):
        ops.uv_backward_warp(torch.zeros(1, 4, 3, 3, device="cuda"), torch.zeros(8, 8, dtype=torch.int64, device="cuda"),
                             torch.zeros(8, 8, 3, device="cuda"), torch.zeros(1, 16, 16, dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        ops.uv_texture_compose(torch.zeros(1, 3, 8, 20, device="cuda"), torch.zeros(1, 1, 8, 20, device="cuda"),
                               torch.zeros(8, 3, 3, device="cuda"), x0=12)
