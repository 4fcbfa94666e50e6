"""Pins the geometry / rasterizer oracle: reference KATs + self-consistency."""
import numpy as np
import pytest
import torch

import oracle
from hoig_b200 import synth
from oracle import geometry_ref as geo


def test_look_at_kat():
    """thirdparty/neural_renderer/tests/test_look_at.py:10-25."""
    eyes = [[1, 0, 1], [0, 0, -10], [-1, 1, 0]]
    answers = [[-np.sqrt(2) / 2, 0, np.sqrt(2) / 2], [1, 0, 10], [0, np.sqrt(2) / 2, 3.0 / 2.0 * np.sqrt(2)]]
    v = torch.tensor([1.0, 0, 0])[None, None, :]
    for e, a in zip(eyes, answers):
        out = geo.look_at(v, np.array(e, np.float32))
        assert np.allclose(out.squeeze().numpy(), np.array(a))


def test_look_at_hogan_eye_is_translation():
    """utils/nmr.py:357 eye = [0,0,-(1/tan30+1)]: identity rotation, z += 2.732."""
    v = torch.randn(2, 7, 3)
    out = geo.look_at(v, [0.0, 0.0, geo.EYE_Z])
    assert torch.allclose(out[..., :2], v[..., :2])
    assert torch.allclose(out[..., 2], v[..., 2] - geo.EYE_Z)


def test_synthetic_mesh_counts():
    sc = synth.make_scene(2, seed=0, obj_faces=2000)
    assert sc.faces_idx.shape == (1538 + 2000, 3)
    assert sc.verts_src.shape == (2, 778 + 7866, 3)
    assert int(sc.faces_idx[:1538].max()) == 777 and int(sc.faces_idx[1538:].min()) == 778
    assert sc.n_verts == int(sc.faces_idx.max()) + 1


def _single_triangle(is_=64):
    # thirdparty/neural_renderer/tests/test_rasterize.py:87-91 geometry, shifted to z=1
    v = np.array([[0.8, 0.8, 1.0], [0.0, -0.5, 1.0], [0.2, -0.4, 1.0]], np.float32)
    return v[None, None]  # (1,1,3,3)


def test_rasterize_single_triangle_properties():
    faces = _single_triangle()
    fim, wim, depth = oracle.rasterize(faces, 64, flip_y=False)
    cov = fim[0] == 0
    assert cov.sum() > 0 and (fim[0][~cov] == -1).all()
    assert np.allclose(wim[0][cov].sum(-1), 1.0, atol=1e-6)
    assert (wim[0][~cov] == 0).all() and (depth[0][~cov] == 100.0).all()
    assert np.allclose(depth[0][cov], 1.0, atol=1e-6)
    # reversed winding is back-face culled (rasterize_cuda_kernel.cu:57,128)
    fim2, _, _ = oracle.rasterize(faces[:, :, ::-1].copy(), 64, flip_y=False)
    assert (fim2 == -1).all()
    # vertical flip == torch.flip(dims=(1,)) (rasterize.py:335-338)
    fim3, wim3, _ = oracle.rasterize(faces, 64, flip_y=True)
    assert np.array_equal(fim3, fim[:, ::-1]) and np.array_equal(wim3, wim[:, ::-1])


def test_rasterize_tie_break_lowest_face_index():
    f = _single_triangle()
    faces = np.concatenate([f, f], 1)  # two coincident faces: strict '<' keeps face 0
    fim, _, _ = oracle.rasterize(faces, 64, flip_y=False)
    assert set(np.unique(fim)) == {-1, 0}


def test_rasterize_degenerate_and_near_far():
    f = _single_triangle()
    degenerate = np.zeros((1, 1, 3, 3), np.float32); degenerate[..., 2] = 1.0
    behind = f.copy(); behind[..., 2] = 0.05        # zp <= near
    faraway = f.copy(); faraway[..., 2] = 150.0     # far <= zp
    fim, wim, _ = oracle.rasterize(np.concatenate([degenerate, behind, faraway], 1), 64, flip_y=False)
    assert (fim == -1).all() and (wim == 0).all()


def test_face_inv_is_inverse():
    sc = synth.make_scene(1, seed=0, obj_faces=500)
    faces = geo.render_faces(sc.cam, sc.verts_src[:, :sc.n_verts], sc.faces_idx).numpy()
    fi = oracle.face_inv(faces, 256).reshape(-1, 3, 3).astype(np.float64)
    p = 0.5 * (faces.reshape(-1, 3, 3)[:, :, :2].astype(np.float64) * 256 + 255)
    P = np.concatenate([p, np.ones((p.shape[0], 3, 1))], 2)       # rows = vertices [x y 1]
    det = np.abs(np.linalg.det(P))
    front = (np.abs(fi).sum((1, 2)) > 0) & (det > 4.0)   # skip sub-pixel slivers: fp32 inverse is ill-conditioned
    assert front.sum() > 100
    # w = face_inv @ [x, y, 1]^T must be the barycentric one-hot at each vertex
    W = np.einsum("fij,fkj->fik", fi[front], P[front])
    assert np.abs(W - np.eye(3)).max() < 2e-2


def test_scene_rasterizes_both_parts_and_T_roundtrip():
    sc = synth.make_scene(2, seed=0, obj_faces=1500)
    nv = sc.n_verts
    fs = geo.render_faces(sc.cam, sc.verts_src[:, :nv], sc.faces_idx)
    fr = geo.render_faces(sc.cam, sc.verts_ref[:, :nv], sc.faces_idx)
    fim_s, wim_s, _ = oracle.rasterize(fs.numpy(), 256)
    fim_r, wim_r, _ = oracle.rasterize(fr.numpy(), 256)
    for fim in (fim_s, fim_r):
        assert (fim >= 0).mean() > 0.02
        assert ((fim >= 0) & (fim < 1538)).any() and (fim >= 1538).any()
    cm = geo.condition_maps(fs, torch.from_numpy(fim_s), torch.from_numpy(fim_r), torch.from_numpy(wim_r),
                            sc.map_fn, sc.sem_full)
    assert cm["src_cond"].shape == (2, 3, 256, 256) and cm["src_seg"].shape == (2, 15, 256, 256)
    T = cm["T"]
    assert ((T == -2).all(-1) == torch.from_numpy(fim_r == -1)).all()
    # identity property: src pose == ref pose -> T is the pixel's own NDC position
    cm2 = geo.condition_maps(fs, torch.from_numpy(fim_s), torch.from_numpy(fim_s), torch.from_numpy(wim_s),
                             sc.map_fn, sc.sem_full)
    ys, xs = np.nonzero(fim_s[0] >= 0)
    t = cm2["T"][0].numpy()[ys, xs]
    # pixel (row y, col x) of the flipped map sits at ndc ((2x+1-256)/256, -(2(255-y)+1-256)/256) in image coords
    assert np.abs(t[:, 0] - (2 * xs + 1 - 256) / 256).max() < 2e-2
    assert np.abs(t[:, 1] - (2 * ys + 1 - 256) / 256).max() < 2e-2


def test_texture_warp_oracle_identity_atlas_resamples_the_image():
    """nmr.py:973-1058 with one source triangle covering the whole image, everything visible, and an atlas whose barycentric
    weights reproduce the atlas pixel centres: the texture is the bilinear resize of the image (grid_sample, align_corners=False)."""
    import torch.nn.functional as F
    size, hu, wu = 32, 24, 40
    g = torch.Generator().manual_seed(2)
    im = torch.rand(2, 3, size, size, generator=g)
    tri = torch.tensor([[-3.0, -3.0], [5.0, -3.0], [-3.0, 5.0]])          # covers [-1,1]^2
    f2v = tri[None, None].repeat(2, 1, 1, 1)                                # (B, F=1, 3, 2), already y-flipped coordinates
    ys = (torch.arange(hu) * 2 + 1) / hu - 1
    xs = (torch.arange(wu) * 2 + 1) / wu - 1
    py, px = torch.meshgrid(ys, xs, indexing="ij")
    w1, w2 = (px + 3) / 8, (py + 3) / 8                                     # p = v0 + w1 (v1 - v0) + w2 (v2 - v0)
    wim_uv = torch.stack([1 - w1 - w2, w1, w2], -1)
    fim_uv = torch.zeros(hu, wu, dtype=torch.int32)
    src_fim = torch.zeros(2, size, size, dtype=torch.int32)
    syn, T, O = geo.texture_backward_warp(im, f2v, src_fim, fim_uv, wim_uv)
    assert O.abs().max().item() == 0
    assert (T[0, ..., 0] - px).abs().max().item() <= 1e-6 and (T[0, ..., 1] - py).abs().max().item() <= 1e-6
    ref = F.interpolate(im, size=(hu, wu), mode="bilinear", align_corners=False)
    assert (syn - ref)[:, :, 2:-2, 2:-2].abs().max().item() <= 1e-5     # the rim differs: zeros padding vs interpolate's clamping
    # nothing visible -> occluded everywhere -> the opened mask paints the texture white
    syn2, _, O2 = geo.texture_backward_warp(im, f2v, src_fim - 1, fim_uv, wim_uv)
    assert O2.min().item() == 1 and (syn2 - 1).abs().max().item() == 0
    # off-mesh pixels of a pose sample to -2 and render as zeros (grid_sample zeros padding)
    fim = torch.full((1, 8, 8), -1, dtype=torch.int32)
    Td = geo.sample_from_texture_dense(fim, torch.ones(1, 8, 8, 3) / 3, torch.zeros(1, 3, 2))
    assert (Td == -2).all()
    assert geo.render_from_texture(syn[:1], fim, torch.ones(1, 8, 8, 3) / 3, torch.zeros(1, 3, 2)).abs().max().item() == 0
