"""Boundaries B2 / B3 on the GPU: the call-compatible stand-ins for the reference's pybind modules and the batched
renderer glue, against the oracle."""
import sys

import numpy as np
import pytest
import torch

import oracle
from hoig_b200 import compat, renderer, synth
from oracle import geometry_ref as geo

pytestmark = pytest.mark.gpu


def test_compat_modules_follow_the_reference_calling_convention():
    be, la, rast = compat.install()
    assert sys.modules["block_extractor_cuda"] is be and sys.modules["local_attn_reshape_cuda"] is la
    g = torch.Generator().manual_seed(0)
    # thirdparty/block_extractor/block_extractor.py:21-26: caller allocates zeros, op fills in place, returns int
    src = torch.rand(2, 6, 14, 10, generator=g).cuda()
    flow = (torch.rand(2, 2, 14, 10, generator=g) * 3.6 - 1.8).cuda()
    out = flow.new(2, 6, 70, 50).zero_()
    assert be.forward(src, flow, out, 5) == 1
    assert torch.equal(out.cpu(), torch.from_numpy(oracle.block_extract(src.cpu().numpy(), flow.cpu().numpy(), 5)))
    # block_extractor.py:34-42: backward adds into caller-provided zero-filled gradients and returns int
    gs, gf = torch.zeros_like(src), torch.zeros_like(flow)
    assert be.backward(src, flow, torch.ones_like(out), gs, gf, 5) == 1
    rs, rf = oracle.block_extract_backward(src.cpu().numpy(), flow.cpu().numpy(), np.ones((2, 6, 70, 50), np.float32), 5)
    assert np.abs(gs.cpu().numpy() - rs).max() <= 1e-3 and np.abs(gf.cpu().numpy() - rf).max() <= 1e-3
    x = torch.rand(4, 9, 14, 10, generator=g).cuda()
    o = x.new(4, 1, 42, 30).zero_()
    assert la.forward(x, o, 3) == 1
    assert torch.equal(o.cpu(), torch.from_numpy(oracle.local_attn_reshape(x.cpu().numpy(), 3)))
    # rasterize.py:50-52 + rasterize_cuda.cpp:70-95: pre-filled outputs, unflipped result, same tensors returned
    sc = synth.make_scene(1, seed=2, obj_faces=800)
    faces = geo.render_faces(sc.cam, sc.verts_src[:, :sc.n_verts], sc.faces_idx).cuda()
    B, F = faces.shape[:2]
    fim = torch.cuda.IntTensor(B, 64, 64).fill_(-1)
    wim = torch.cuda.FloatTensor(B, 64, 64, 3).fill_(0.0)
    depth = torch.cuda.FloatTensor(B, 64, 64).fill_(100.0)
    finv = torch.zeros_like(faces)
    ret = rast.forward_face_index_map(faces, fim, wim, depth, torch.cuda.FloatTensor(1).fill_(0), finv, 64, 0.1, 100.0, 0, 0, 0)
    assert ret[0] is fim and ret[1] is wim and ret[2] is depth
    fo, wo, do = oracle.rasterize(faces.cpu().numpy(), 64, flip_y=False)
    assert np.array_equal(fim.cpu().numpy(), fo) and np.array_equal(wim.cpu().numpy(), wo) and np.array_equal(depth.cpu().numpy(), do)
    assert np.array_equal(finv.cpu().numpy().reshape(-1, 9), oracle.face_inv(faces.cpu().numpy(), 64))
    with pytest.raises(RuntimeError):
        rast.forward_face_index_map(faces.cpu(), fim, wim, depth, depth, finv, 64, 0.1, 100.0, 0, 0, 0)
    # the renderer-level API returns the flipped maps like rasterize.py:543-571
    f2, w2 = renderer.rasterize_face_index_map_and_weight_map(faces, 64, False)
    assert torch.equal(f2, torch.flip(fim, dims=(1,))) and torch.equal(w2, torch.flip(wim, dims=(1,)))


def test_batched_condition_inputs_match_oracle():
    """models/trainer.py:63-145 batched: rasterize both poses, gather tables, masks, T_hand, generator inputs."""
    B = 3
    sc = synth.make_scene(B, seed=4, obj_faces=3000)
    nv = sc.n_verts
    cam, fidx = sc.cam.cuda(), sc.faces_idx.cuda()
    faces_s, fim_s, wim_s = renderer.render_fim_wim_batched(cam, sc.verts_src[:, :nv].contiguous().cuda(), fidx)
    faces_r, fim_r, wim_r = renderer.render_fim_wim_batched(cam, sc.verts_ref[:, :nv].contiguous().cuda(), fidx)
    src_img = torch.rand(B, 3, 256, 256).cuda() * 2 - 1
    inp, masks = renderer.condition_inputs(src_img, faces_s, fim_s, fim_r, wim_r, sc.map_fn.cuda(), sc.sem_full.cuda())
    # oracle on the same faces (the rasterizer itself is gated bit-exact elsewhere)
    fo_s, wo_s, _ = oracle.rasterize(faces_s.cpu().numpy(), 256)
    fo_r, wo_r, _ = oracle.rasterize(faces_r.cpu().numpy(), 256)
    assert np.array_equal(fim_s.cpu().numpy(), fo_s) and np.array_equal(fim_r.cpu().numpy(), fo_r)
    cm = geo.condition_maps(faces_s.cpu(), torch.from_numpy(fo_s), torch.from_numpy(fo_r), torch.from_numpy(wo_r), sc.map_fn, sc.sem_full)
    assert torch.equal(masks["src_mask_hand"].cpu(), cm["src_mask_hand"]) and torch.equal(masks["ref_mask_bg"].cpu(), cm["ref_mask_bg"])
    assert (inp["T"].cpu() - cm["T_hand"]).abs().max().item() <= 1e-6
    assert torch.equal(inp["src_hand_conds"].cpu(), cm["src_cond_hand"])
    assert torch.equal(inp["tsf_obj_conds"].cpu(), torch.cat([cm["ref_cond_obj"], cm["ref_seg"][:, 6:]], 1))
    bgm = cm["src_bg_mask15"]
    assert torch.allclose(inp["bg_inputs"].cpu(), torch.cat([src_img.cpu() * bgm, bgm], 1))
    assert inp["src_obj_inputs"].shape == (B, 3, 256, 256) and inp["src_obj_conds"].shape == (B, 12, 256, 256)
    assert inp["src_hand_inputs"].shape == (B, 3, 256, 256) and inp["T"].shape == (B, 256, 256, 2)


def test_fused_condition_inputs_equal_the_stepwise_glue():
    """Row N1: ``hoig_condition_inputs`` (one launch) reproduces ``renderer.condition_inputs`` (C-ABI kernels + torch glue, gated
    against the oracle above) bit for bit, with distinct re-rendered images on both sides."""
    B = 3
    sc = synth.make_scene(B, seed=7, obj_faces=3000)
    nv = sc.n_verts
    cam, fidx = sc.cam.cuda(), sc.faces_idx.cuda()
    faces_s, fim_s, wim_s = renderer.render_fim_wim_batched(cam, sc.verts_src[:, :nv].contiguous().cuda(), fidx)
    faces_r, fim_r, wim_r = renderer.render_fim_wim_batched(cam, sc.verts_ref[:, :nv].contiguous().cuda(), fidx)
    g = torch.Generator().manual_seed(3)
    src_img, r_src, r_ref = [(torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).cuda() for _ in range(3)]
    args = (src_img, faces_s, fim_s, fim_r, wim_r, sc.map_fn.cuda(), sc.sem_full.cuda(), r_src, r_ref)
    a, ma = renderer.condition_inputs(*args)
    b, mb = renderer.condition_inputs_fused(*args)
    torch.cuda.synchronize()
    assert set(a) == set(b) and set(ma) == set(mb)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    for k in ma:
        assert torch.equal(ma[k], mb[k]), k
    assert 0 < mb["ref_mask_hand"].mean().item() < 1 and (b["T"][..., 0] > -1.5).any()


def test_meshes_to_image_end_to_end():
    """The whole hot path on device: meshes + source image -> condition stage (R0-R8, N1) -> generator -> composite.  Checks the
    plumbing between the stages (shapes, dtypes, value ranges, no NaN), each stage being gated against its oracle elsewhere."""
    from hoig_b200.generator import composite, create
    B = 2
    sc = synth.make_scene(B, seed=11, obj_faces=3000)
    nv = sc.n_verts
    cam, fidx = sc.cam.cuda(), sc.faces_idx.cuda()
    faces_s, fim_s, wim_s = renderer.render_fim_wim_batched(cam, sc.verts_src[:, :nv].contiguous().cuda(), fidx)
    _, fim_r, wim_r = renderer.render_fim_wim_batched(cam, sc.verts_ref[:, :nv].contiguous().cuda(), fidx)
    g = torch.Generator().manual_seed(0)
    src_img = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).cuda()
    F = fidx.shape[0]
    fim_uv = torch.randint(-1, F, (256, 640), generator=g, dtype=torch.int32).cuda()
    w = torch.rand(256, 640, 3, generator=g) + 0.05
    wim_uv = (w / w.sum(-1, keepdim=True)).cuda()
    uv_coord = (torch.rand(F, 3, 2, generator=g) * 2 - 1).cuda()
    tex = renderer.texture_backward_warp(src_img, faces_s, fim_s, fim_uv, wim_uv, torch.rand(256, 256, 3, generator=g).cuda())
    r_src = renderer.render_from_texture(tex, fim_s, wim_s, uv_coord)
    r_ref = renderer.render_from_texture(tex, fim_r, wim_r, uv_coord)
    inputs, masks = renderer.condition_inputs_fused(src_img, faces_s, fim_s, fim_r, wim_r, sc.map_fn.cuda(), sc.sem_full.cuda(), r_src, r_ref)
    net = create("generator_spade_attn", dtype=torch.float16, bg_dim=8, img_dim=3, obj_dim=3, img_cond_dim=3, obj_cond_dim=12,
                 conv_dim=16, repeat_num=2)
    net.init_weights()
    net = net.cuda().eval()
    outs = net(**inputs, src_armask=torch.zeros(B, 1, 256, 256).cuda(), tsf_armask=torch.zeros(B, 1, 256, 256).cuda())
    img = composite(outs[1], outs[6], outs[7], outs[8], outs[9])
    torch.cuda.synchronize()
    assert img.shape == (B, 3, 256, 256) and torch.isfinite(img).all()
    assert img.abs().max().item() <= 1.0 + 1e-5            # tanh images blended by sigmoid masks
    assert all(torch.isfinite(o).all() for o in outs)
