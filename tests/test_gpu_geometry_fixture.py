"""Stage R (R0-R8, rows N1/N2) CUDA path through the C ABI against the fixture produced by the UNMODIFIED reference
(tests/golden/geometry_stage_r.npz: HandRecoveryFlow.forward + the MANORenderer methods + util.morph of /root/reference/HOIG_HOv3
run on the seeded scene of tests/geometry_inputs.py by tests/golden/make_golden_geometry.py)."""
import os

import numpy as np
import pytest
import torch

from hoig_b200 import ops, renderer

from . import geometry_inputs as gi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx(golden_dir):
    return np.load(os.path.join(golden_dir, "geometry_stage_r.npz"))


@pytest.fixture(scope="module")
def inp():
    d = gi.scene_and_tables()
    sc = d["scene"]
    got = gi.checksum(sc.faces_idx, sc.verts_src, sc.verts_ref, sc.cam, d["map_fn"], d["sem_full"], d["fim_uv"], d["wim_uv"],
                      d["coord"], d["src_img"], d["obj_tex"])
    return d, got


def _c(a):
    return torch.from_numpy(np.asarray(a)).cuda()


def _wim(fx, tag):
    fim = fx[f"fim_{tag}"]
    wim = np.zeros(fim.shape + (3,), np.float32)
    wim[fim != -1] = fx[f"wim_{tag}_covered"]
    return _c(wim)


def test_inputs_reproduce(fx, inp):
    assert inp[1] == str(fx["input_checksum"])


@pytest.mark.parametrize("tag", ["src", "ref"])
def test_project_and_rasterize_vs_reference(fx, inp, tag):
    """R0-R3: hoig_project_faces within 2e-6 of the reference's projected faces; the rasterizer on the reference's faces gives the
    reference's fim / wim bit for bit; and the whole mesh -> map path differs from it on a vanishing fraction of edge pixels only."""
    d, _ = inp
    sc = d["scene"]
    verts = (sc.verts_src if tag == "src" else sc.verts_ref).cuda()
    faces, fim, wim = renderer.render_fim_wim_batched(sc.cam.cuda(), verts, sc.faces_idx.cuda())
    assert (faces.cpu() - torch.from_numpy(fx[f"faces_{tag}"])).abs().max().item() <= 2e-6
    fim_r, wim_r = ops.rasterize(_c(fx[f"faces_{tag}"]), 256)
    assert np.array_equal(fim_r.cpu().numpy(), fx[f"fim_{tag}"])
    assert torch.equal(wim_r, _wim(fx, tag))
    assert (fim.cpu().numpy() != fx[f"fim_{tag}"]).mean() <= 1e-4


def test_condition_inputs_vs_reference(fx, inp):
    """R4-R7 + N1: hoig_condition_inputs (one launch) and the stepwise kernels against HandRecoveryFlow.forward's outputs."""
    d, _ = inp
    src_img = d["src_img"].cuda()
    faces_src, fim_s, fim_r, wim_r = _c(fx["faces_src"]), _c(fx["fim_src"]), _c(fx["fim_ref"]), _wim(fx, "ref")
    map_fn, sem = d["map_fn"].cuda(), d["sem_full"].cuda()
    # texture stage first: its outputs feed the rgb inputs
    tex = renderer.texture_backward_warp(src_img, faces_src, fim_s, d["fim_uv"].cuda(), d["wim_uv"].cuda(), d["obj_tex"].cuda(), 384)
    s = gi.STRIDE
    assert (tex[..., ::s, ::s].cpu() - torch.from_numpy(fx["texture_s"])).abs().max().item() <= 2e-6
    assert abs(tex.double().sum().item() - fx["texture_sum"][0]) <= 2e-6 * tex.numel()
    coord = d["coord"].cuda()
    T_src = renderer.sample_from_texture_dense(fim_s, _wim(fx, "src"), coord)
    T_ref = renderer.sample_from_texture_dense(fim_r, wim_r, coord)
    assert torch.equal(T_src.cpu(), torch.from_numpy(fx["T_tex_src"]))
    assert torch.equal(T_ref.cpu(), torch.from_numpy(fx["T_tex_ref"]))
    r_src = renderer.render_from_texture(tex, fim_s, _wim(fx, "src"), coord)
    r_ref = renderer.render_from_texture(tex, fim_r, wim_r, coord)
    for fused in (True, False):
        fn = renderer.condition_inputs_fused if fused else renderer.condition_inputs
        out, masks = fn(src_img, faces_src, fim_s, fim_r, wim_r, map_fn, sem, r_src, r_ref)
        torch.cuda.synchronize()

        def eq(t, key):
            assert torch.equal(t.cpu(), torch.from_numpy(fx[key]).to(t.dtype)), (key, fused)

        eq(masks["src_mask_bg"], "out_src_crop_mask_bg"); eq(masks["ref_mask_bg"], "out_ref_crop_mask_bg")
        eq(masks["src_mask_hand"], "out_src_crop_mask_hand"); eq(masks["ref_mask_hand"], "out_ref_crop_mask_hand")
        eq(out["bg_inputs"][:, 3:], "out_input_G_src_bg_mask")
        eq(out["src_obj_conds"][:, :3], "out_input_G_src_obj_cond"); eq(out["src_obj_conds"][:, 3:], "out_input_G_src_obj_seg")
        eq(out["tsf_obj_conds"][:, :3], "out_input_G_tsf_obj_cond"); eq(out["tsf_obj_conds"][:, 3:], "out_input_G_tsf_obj_seg")
        eq(out["src_hand_conds"], "out_input_G_src_hand_cond"); eq(out["tsf_hand_conds"], "out_input_G_ref_hand_cond")
        assert (out["T"].cpu() - torch.from_numpy(fx["out_T_hand"])).abs().max().item() <= 1e-6
        for k, key in (("bg_inputs", "out_input_G_src_bg_rgb"), ("src_obj_inputs", "out_input_G_src_obj_rgb"),
                       ("tsf_obj_inputs", "out_input_G_tsf_obj_rgb"), ("src_hand_inputs", "out_input_G_src_hand_rgb"),
                       ("tsf_hand_inputs", "out_input_G_ref_hand_rgb")):
            t = out[k][:, :3]
            assert (t[..., ::s, ::s].cpu() - torch.from_numpy(fx[key + "_s"])).abs().max().item() <= 2e-6, (k, fused)
            assert abs(t.double().sum().item() - fx[key + "_sum"][0]) <= 2e-6 * t.numel(), (k, fused)


def test_stepwise_kernels_vs_reference(fx, inp):
    """R4 (encode_fim / encode_sem), R6 (morph) and R7 (cal_bc_transform) one at a time."""
    d, _ = inp
    for tag in ("src", "ref"):
        fim = _c(fx[f"fim_{tag}"])
        cond, seg, not_hand = ops.condition_maps(fim, d["map_fn"].cuda(), d["sem_full"].cuda(), renderer.N_HAND_FACES)
        assert torch.equal(cond.cpu(), torch.from_numpy(fx[f"cond_{tag}"]))
        sem = torch.from_numpy(fx[f"sem_{tag}"]).float()
        assert torch.equal(seg.cpu(), torch.cat([(sem == i).float() for i in range(1, 16)], 1))
        key = "out_src_crop_mask_hand" if tag == "src" else "out_ref_crop_mask_hand"
        assert torch.equal(ops.erode(not_hand, 3).cpu(), torch.from_numpy(fx[key]).float())
    T = ops.bc_transform(_c(fx["faces_src"]), _c(fx["fim_ref"]), _wim(fx, "ref"))
    assert (T.cpu() - torch.from_numpy(fx["bc_T"])).abs().max().item() <= 1e-6


def test_hand_recovery_flow_module_end_to_end_vs_reference(fx, inp):
    """Row N1: ``HandRecoveryFlowB200`` from MESHES (own projection + rasterization + texture warp + conditions) against the outputs of
    the unmodified ``HandRecoveryFlow.forward``.  The projection differs from torch's in the last ulp (<= 2e-6), which moves a handful
    of silhouette-edge pixels; everything else must agree."""
    d, _ = inp
    sc = d["scene"]
    flow = renderer.HandRecoveryFlowB200(sc.faces_idx, d["map_fn"], d["sem_full"], d["fim_uv"], d["wim_uv"], d["coord"], d["obj_tex"]).cuda()
    kw, masks = flow(d["src_img"].cuda(), sc.verts_src.cuda(), sc.verts_ref.cuda(), sc.cam.cuda())
    torch.cuda.synchronize()
    assert set(kw) == {"bg_inputs", "src_obj_inputs", "src_obj_conds", "src_hand_inputs", "src_hand_conds", "tsf_obj_inputs",
                       "tsf_obj_conds", "tsf_hand_inputs", "tsf_hand_conds", "T"}

    def frac_diff(t, key):
        ref = torch.from_numpy(fx[key]).to(t.dtype)
        return (t.cpu() != ref).float().mean().item()

    assert frac_diff(masks["src_mask_bg"], "out_src_crop_mask_bg") <= 2e-4
    assert frac_diff(masks["ref_mask_hand"], "out_ref_crop_mask_hand") <= 2e-4
    assert frac_diff(kw["bg_inputs"][:, 3:], "out_input_G_src_bg_mask") <= 2e-4
    assert frac_diff(kw["src_obj_conds"][:, 3:], "out_input_G_src_obj_seg") <= 2e-4
    Tref = torch.from_numpy(fx["out_T_hand"])
    both = (kw["T"].cpu()[..., 0] > -1.5) & (Tref[..., 0] > -1.5)
    assert both.float().mean().item() > 0.01
    assert (kw["T"].cpu() - Tref)[both].abs().max().item() <= 5e-3           # sub-pixel agreement where both are on the hand
    assert ((kw["T"].cpu()[..., 0] > -1.5) != (Tref[..., 0] > -1.5)).float().mean().item() <= 2e-4
    s = gi.STRIDE
    a, b = kw["src_hand_inputs"][..., ::s, ::s].cpu(), torch.from_numpy(fx["out_input_G_src_hand_rgb_s"])
    assert ((a - b).abs() > 1e-5).float().mean().item() <= 1e-3
