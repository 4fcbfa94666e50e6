"""GPU parity of stage R (rasterizer + condition maps) through the C ABI.

Bit-exact face-index maps against (a) the C oracle and (b), when oracle/_ref was
built, the reference's own kernel compiled with the same nvcc, on the same GPU.
"""
import os

import numpy as np
import pytest
import torch

import oracle
from hoig_b200 import ops, synth
from oracle import geometry_ref as geo

pytestmark = pytest.mark.gpu
from oracle.ref_kernels import load_ref as _load_ref, ref_rasterize as _ref_rasterize  # noqa: E402


def _scene_faces(B, seed, obj_faces):
    sc = synth.make_scene(B, seed=seed, obj_faces=obj_faces)
    nv = sc.n_verts
    fs = geo.render_faces(sc.cam, sc.verts_src[:, :nv], sc.faces_idx)
    fr = geo.render_faces(sc.cam, sc.verts_ref[:, :nv], sc.faces_idx)
    return sc, torch.cat([fs, fr], 0).contiguous()


def _compare(faces_cpu, is_, tag):
    fim_o, wim_o, dep_o = oracle.rasterize(faces_cpu.numpy(), is_)
    fim, wim, dep = ops.rasterize(faces_cpu.cuda(), is_, return_depth=True)
    torch.cuda.synchronize()
    # the binning pre-pass (workspace) and the scan-all-faces path must agree bit for bit
    f2, w2, d2 = ops.rasterize(faces_cpu.cuda(), is_, return_depth=True, use_workspace=False)
    assert torch.equal(f2, fim) and torch.equal(w2, wim) and torch.equal(d2, dep)
    fim, wim, dep = fim.cpu().numpy(), wim.cpu().numpy(), dep.cpu().numpy()
    nbad = int((fim != fim_o).sum())
    print(f"[{tag}] covered={int((fim_o >= 0).sum())} fim mismatches={nbad} "
          f"wim maxabs={np.abs(wim - wim_o).max():.3e} depth maxabs={np.abs(dep - dep_o).max():.3e}")
    assert nbad == 0
    assert np.array_equal(wim, wim_o)      # same op sequence -> identical bits (0 == -0 compares equal)
    assert np.array_equal(dep, dep_o)
    return fim, wim, dep


@pytest.mark.parametrize("is_,obj_faces", [(256, 3000), (128, 800), (64, 400)])
def test_fim_bit_exact_vs_c_oracle(is_, obj_faces):
    _, faces = _scene_faces(2, seed=is_, obj_faces=obj_faces)
    fim, _, _ = _compare(faces, is_, f"scene is={is_}")
    assert (fim >= 0).mean() > 0.01


def test_fim_bit_exact_full_face_count():
    _, faces = _scene_faces(1, seed=7, obj_faces=12238)   # F = 13776 as in utils/nmr.py:877
    assert faces.shape[1] == 13776
    _compare(faces, 256, "F=13776")


def test_edge_cases_match_oracle():
    g = torch.Generator().manual_seed(0)
    tri = torch.tensor([[0.8, 0.8, 1.0], [0.0, -0.5, 1.0], [0.2, -0.4, 1.0]])
    faces = [tri, tri.clone(),                      # coincident: lowest index wins
             tri[[0, 2, 1]],                        # back-facing
             torch.zeros(3, 3) + torch.tensor([0.1, 0.1, 1.0]),   # point-degenerate
             torch.tensor([[-0.9, -0.9, 2.0], [0.9, 0.9, 2.0], [0.0, 0.0, 2.0]]),  # collinear
             torch.tensor([[-0.5, 0.3, 0.05], [0.5, 0.3, 0.05], [0.0, 0.9, 0.05]]),  # in front of near
             torch.tensor([[-0.5, 0.3, 150.0], [0.5, 0.3, 150.0], [0.0, 0.9, 150.0]]),  # beyond far
             torch.tensor([[float("nan"), 0.0, 1.0], [0.5, 0.3, 1.0], [0.0, 0.9, 1.0]]),
             torch.tensor([[float("inf"), 0.0, 1.0], [0.5, 0.3, 1.0], [0.0, 0.9, 1.0]]),
             torch.tensor([[1e20, -1e20, 1.0], [-1e20, -1e20, 1.0], [0.0, 1e20, 1.0]]),   # 'wild' path, covers the screen
             torch.tensor([[-3.0, -3.0, 5.0], [3.0, -3.0, 5.0], [0.0, 3.0, 5.0]])]       # large, partially off-screen
    for _ in range(40):                               # random needles and slivers
        a = torch.rand(2, generator=g) * 2 - 1
        d = torch.rand(2, generator=g) * 2 - 1
        eps = (torch.rand(1, generator=g).item() - 0.5) * 1e-4
        p0, p1 = a, a + d
        p2 = a + 0.5 * d + eps * torch.tensor([-d[1], d[0]])
        z = 1.0 + torch.rand(3, generator=g)
        f = torch.cat([torch.stack([p0, p1, p2]), z[:, None]], 1)
        faces += [f, f[[0, 2, 1]]]
    faces = torch.stack(faces)[None].contiguous()
    for is_ in (64, 256):
        _compare(faces, is_, f"edge cases is={is_}")


def test_empty_inputs():
    fim, wim = ops.rasterize(torch.zeros(0, 5, 3, 3, device="cuda"), 64)
    assert fim.shape == (0, 64, 64)
    fim, wim = ops.rasterize(torch.zeros(2, 0, 3, 3, device="cuda"), 64)
    assert (fim == -1).all() and (wim == 0).all()


def test_fim_bit_exact_vs_reference_kernel():
    mod = _load_ref("ref_rasterize_cuda")
    if mod is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    _, faces = _scene_faces(2, seed=11, obj_faces=12238)
    faces = faces.cuda()
    for is_ in (256, 64):
        rf, rw, rd, rfinv = _ref_rasterize(mod, faces, is_)
        fim, wim, dep = ops.rasterize(faces, is_, return_depth=True)
        nbad = int((fim != rf).sum())
        print(f"[vs reference kernel is={is_}] covered={int((rf >= 0).sum())} fim mismatches={nbad} "
              f"wim maxabs={(wim - rw).abs().max().item():.3e} depth maxabs={(dep - rd).abs().max().item():.3e}")
        assert nbad == 0
        assert torch.equal(wim, rw) and torch.equal(dep, rd)
        finv = ops.face_inv(faces, is_)
        assert torch.equal(finv.view_as(rfinv), rfinv)
        # and the C oracle agrees with the reference kernel too (pins the restatement)
        fo, wo, do = oracle.rasterize(faces.cpu().numpy(), is_)
        assert np.array_equal(fo, rf.cpu().numpy()) and np.array_equal(wo, rw.cpu().numpy())


def test_round_trip_property_at_scale():
    """Size-independent check for large batches: identical src/ref pose => T maps every covered
    pixel to its own NDC centre (encode -> correspond -> identity)."""
    B = 64
    sc = synth.make_scene(B, seed=5, obj_faces=12238)
    faces = ops.project_faces(sc.verts_src[:, :sc.n_verts].contiguous().cuda(), sc.cam.cuda(), sc.faces_idx.cuda(), geo.EYE_Z)
    fim, wim = ops.rasterize(faces, 256)
    T = ops.bc_transform(faces, fim, wim)
    cov = fim >= 0
    assert cov.float().mean().item() > 0.01
    ys, xs = torch.meshgrid(torch.arange(256, device="cuda"), torch.arange(256, device="cuda"), indexing="ij")
    ex = ((2 * xs + 1 - 256) / 256.0)[None].expand(B, -1, -1)
    ey = ((2 * ys + 1 - 256) / 256.0)[None].expand(B, -1, -1)
    assert (T[..., 0] - ex)[cov].abs().max().item() < 2e-2
    assert (T[..., 1] - ey)[cov].abs().max().item() < 2e-2
    assert (T[~cov] == -2).all()
    w = wim[cov]
    assert (w.sum(-1) - 1).abs().max().item() < 1e-5 and (w >= 0).all()


def test_projection_and_condition_maps_vs_oracle():
    sc = synth.make_scene(2, seed=3, obj_faces=2000)
    nv = sc.n_verts
    faces_o = geo.render_faces(sc.cam, sc.verts_src[:, :nv], sc.faces_idx)
    faces = ops.project_faces(sc.verts_src[:, :nv].contiguous().cuda(), sc.cam.cuda(), sc.faces_idx.cuda(), geo.EYE_Z)
    err = (faces.cpu() - faces_o).abs().max().item()
    print("project_faces max abs err", err)
    assert err <= 2e-6
    faces_r = geo.render_faces(sc.cam, sc.verts_ref[:, :nv], sc.faces_idx)
    fim_s, wim_s, _ = oracle.rasterize(faces_o.numpy(), 256)
    fim_r, wim_r, _ = oracle.rasterize(faces_r.numpy(), 256)
    cm = geo.condition_maps(faces_o, torch.from_numpy(fim_s), torch.from_numpy(fim_r), torch.from_numpy(wim_r), sc.map_fn, sc.sem_full)
    cond, seg, not_hand = ops.condition_maps(torch.from_numpy(fim_s).cuda(), sc.map_fn.cuda(), sc.sem_full.cuda(), 1538)
    assert torch.equal(cond.cpu(), cm["src_cond"]) and torch.equal(seg.cpu(), cm["src_seg"])
    assert torch.equal(ops.erode(not_hand, 3).cpu(), cm["src_mask_hand"])
    assert torch.equal(ops.erode(cond[:, -1:].contiguous(), 15).cpu(), cm["src_bg_mask15"])
    T = ops.bc_transform(faces_o.cuda(), torch.from_numpy(fim_r).cuda(), torch.from_numpy(wim_r).cuda())
    assert (T.cpu() - cm["T"]).abs().max().item() <= 1e-6
