"""Restatement of the HandRecoveryFlow geometry glue -- TEST INFRASTRUCTURE ONLY.

Torch-CPU fp32 restatement of stages R0 and R4-R7 of SURVEY.md section 8a
(paths relative to /root/reference/HOIG_HOv3):

* R0  ``utils/nmr.py:109-140`` orthographic_proj_withz_idrot, ``:506`` y flip,
      ``thirdparty/neural_renderer/neural_renderer/look_at.py:6-62``,
      ``.../vertices_to_faces.py:4-22``
* R4  ``utils/nmr.py:567-595`` encode_fim / encode_sem (fim=-1 hits the last row)
* R5  ``models/trainer.py:71-72,109-136`` one-hot seg, hand/bg masks, cond split
* R6  ``utils/util.py:142-158`` morph (erode)
* R7  ``utils/nmr.py:874-968`` cal_bc_transform (T only; O is discarded by the
      caller, trainer.py:80) and ``trainer.py:81`` T_hand

Pinned to the reference itself: tests/golden/geometry_stage_r.npz is produced by running the UNMODIFIED
``HandRecoveryFlow.forward`` / ``MANORenderer`` methods / ``util.morph`` from /root/reference
(tests/golden/make_golden_geometry.py), and tests/test_geometry_fixture_cpu.py holds every function below
to it (bit-equal for T, masks, encoded maps; <= 1e-6 for projected faces and textures); plus the reference's
look_at KAT (tests/test_look_at.py:10-19) in tests/test_oracle_geometry.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

EYE_Z = -(1.0 / math.tan(math.radians(30.0)) + 1.0)  # utils/nmr.py:357


def project(verts: torch.Tensor, cam: torch.Tensor, offset_z: float = 0.0) -> torch.Tensor:
    """utils/nmr.py:109-140.  verts (B,V,3) OpenGL coords, cam (B,15)."""
    bs = cam.shape[0]
    cam_mat = cam[:, 0:9].reshape(bs, 3, 3)
    trans = cam[:, 9:].reshape(bs, 2, 3)
    cc = torch.tensor([[1.0, 0, 0], [0, -1.0, 0], [0, 0, -1.0]], dtype=torch.float32)[None].repeat(bs, 1, 1)
    p = torch.einsum("ijk,imk->ijm", [verts, cc])
    pp = torch.einsum("ijk,imk->ijm", [p, cam_mat])
    z = pp[:, :, 2]
    xy = torch.stack([pp[:, :, 0] / z, pp[:, :, 1] / z], 2)
    xy1 = torch.cat([xy, torch.ones_like(xy)[:, :, 1:2]], 2)
    xyt = torch.einsum("ijk,imk->ijm", [trans, xy1]).permute(0, 2, 1)
    xyt = xyt / 255.0 * 2 - 1
    return torch.cat((xyt, p[:, :, 2:3] + offset_z), 2)


def look_at(vertices: torch.Tensor, eye, at=(0, 0, 0), up=(0, 1, 0)) -> torch.Tensor:
    """thirdparty/neural_renderer/neural_renderer/look_at.py:6-62."""
    eye = torch.as_tensor(np.asarray(eye, np.float32))
    at = torch.as_tensor(np.asarray(at, np.float32))
    up = torch.as_tensor(np.asarray(up, np.float32))
    B = vertices.shape[0]
    eye, at, up = (t[None].repeat(B, 1) if t.ndim == 1 else t for t in (eye, at, up))
    z = F.normalize(at - eye, eps=1e-5)
    x = F.normalize(torch.cross(up, z, dim=1), eps=1e-5)
    y = F.normalize(torch.cross(z, x, dim=1), eps=1e-5)
    r = torch.stack([x, y, z], 1)
    return torch.matmul(vertices - eye[:, None, :], r.transpose(1, 2))


def vertices_to_faces(vertices: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """thirdparty/neural_renderer/neural_renderer/vertices_to_faces.py:4-22."""
    B, nv = vertices.shape[:2]
    idx = faces.long() + (torch.arange(B) * nv)[:, None, None]
    return vertices.reshape(B * nv, 3)[idx]


def render_faces(cam: torch.Tensor, verts: torch.Tensor, faces_idx: torch.Tensor) -> torch.Tensor:
    """utils/nmr.py:496-511 up to the rasterizer call: (B,F,3,3) camera-space faces."""
    pv = project(verts, cam)
    pv = pv.clone()
    pv[:, :, 1] *= -1
    v = look_at(pv, [0.0, 0.0, EYE_Z])
    if faces_idx.ndim == 2:
        faces_idx = faces_idx[None].repeat(cam.shape[0], 1, 1)
    return vertices_to_faces(v, faces_idx)


def encode(fim: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """utils/nmr.py:567-595: ``table[fim.long()]`` then NCHW; -1 wraps to the last (bg) row."""
    return table[fim.long()].permute(0, 3, 1, 2)


def erode(mask: torch.Tensor, ks: int) -> torch.Tensor:
    """utils/util.py:142-153 (mode='erode')."""
    pad = ks // 2
    m = F.pad(mask, [pad] * 4, value=1.0)
    out = F.conv2d(m, torch.ones(1, 1, ks, ks))
    return (out == ks * ks).float()


def bc_transform(src_f2pts: torch.Tensor, dst_fim: torch.Tensor, dst_wim: torch.Tensor) -> torch.Tensor:
    """utils/nmr.py:874-925: T = sum_k src_f2pts[fim][k] * wim[k] where fim != -1, else -2."""
    B, H, W = dst_fim.shape
    T = -2 * torch.ones(B, H * W, 2)
    for i in range(B):
        f = dst_fim[i].long().reshape(-1)
        w = dst_wim[i].reshape(-1, 3)
        m = f != -1
        T[i, m] = (src_f2pts[i][f[m]] * w[m][:, :, None]).sum(dim=1)
    return T.view(B, H, W, 2)


def condition_maps(faces_src, fim_src, fim_ref, wim_ref, map_fn, sem_full, n_hand_faces: int = 1538):
    """models/trainer.py:66-81,109-124 for a batch (the reference loops per sample).

    Returns a dict with cond/seg/masks for src and ref plus T_hand.
    """
    f2v = faces_src[:, :, :, 0:2].clone()
    f2v[:, :, :, 1] *= -1
    out = {}
    for tag, fim in (("src", fim_src), ("ref", fim_ref)):
        cond = encode(fim, map_fn)
        sem = encode(fim, sem_full)
        out[f"{tag}_cond"] = cond
        out[f"{tag}_seg"] = torch.cat([(sem == i).float() for i in range(1, 16)], 1)
        out[f"{tag}_mask_hand"] = erode(1 - ((fim != -1) & (fim < n_hand_faces))[:, None].float(), 3)
        out[f"{tag}_mask_bg"] = erode(cond[:, -1:], 3)
        out[f"{tag}_bg_mask15"] = erode(cond[:, -1:], 15)
        hm = (cond[:, :1] < 1.5).float()
        out[f"{tag}_cond_hand"] = torch.cat([hm * cond[:, :2], cond[:, 2:] + 1 - hm], 1)
        om = (cond[:, :1] > 1.5).float()
        out[f"{tag}_cond_obj"] = torch.cat([om * cond[:, :2], cond[:, 2:] + 1 - om], 1)
    T = bc_transform(f2v, fim_ref, wim_ref)
    mh = out["ref_mask_hand"][:, 0][:, :, :, None]
    out["T"] = T
    out["T_hand"] = T * (mh == 0) + (-2) * torch.ones_like(T) * (mh == 1)
    return out


# ------------------------------------------------------------------ stage R8: UV-texture warp (test infrastructure)
def texture_backward_warp(im, src_f2verts, src_fims, fim_uv, wim_uv, obj_tex=None, x0=384):
    """utils/nmr.py:973-1058 restated for CPU tensors, same loops and op order.  src_f2verts (B,F,3,2) are the projected source
    face vertices AFTER trainer.py:67-68 (xy only, y negated); fim_uv (Hu,Wu) long/int, wim_uv (Hu,Wu,3); src_fims (B,is,is)."""
    import torch.nn.functional as F
    bs = src_f2verts.shape[0]
    hu, wu = fim_uv.shape
    size = src_fims.shape[-1]
    T = -2 * torch.ones((bs, hu * wu, 2), dtype=torch.float32)
    O = torch.zeros((bs, hu * wu, 1), dtype=torch.float32)
    for i in range(bs):
        to_fim = fim_uv.long().reshape(-1)
        to_wim = wim_uv.reshape(-1, 3)
        exist = to_fim != -1
        face_idx = to_fim[exist]
        weights = to_wim[exist]
        smpl_T = (src_f2verts[i][face_idx] * weights[:, :, None]).sum(dim=1)
        T[i, exist] = smpl_T
        from_fim = src_fims[i].long().reshape(-1)
        t11 = ((smpl_T + 1) / 2.0 * float(size - 1)).long().clamp(0, size - 1)
        visible = torch.zeros_like(face_idx, dtype=torch.bool)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                t = (t11 + torch.tensor([dx, dy])).clamp(0, size - 1)
                visible |= from_fim[t[:, 1] * size + t[:, 0]] == face_idx
        O[i, exist, 0] = 1 - visible.float()
    T = T.view(bs, hu, wu, 2)
    O = O.view(bs, hu, wu, 1).permute(0, 3, 1, 2)
    syn = F.grid_sample(im, T, align_corners=False)
    O = erode(O, 3)
    O = 1 - erode(1 - O, 3)
    syn = syn * (1 - O) + 1.0 * torch.ones_like(syn) * O
    if obj_tex is not None:
        syn[:, :, :, x0:] = obj_tex.permute(2, 0, 1)[None]
    return syn, T, O


def sample_from_texture_dense(fim, wim, faces_uv_coord):
    """utils/nmr.py:1068-1100."""
    bs, h, w = fim.shape
    T = -2 * torch.ones((bs, h * w, 2), dtype=torch.float32)
    for i in range(bs):
        f = fim[i].long().reshape(-1)
        exist = f != -1
        T[i, exist] = (faces_uv_coord[f[exist]] * wim[i].reshape(-1, 3)[exist][:, :, None]).sum(dim=1)
    return T.view(bs, h, w, 2)


def render_from_texture(texture, fim, wim, faces_uv_coord):
    """models/trainer.py:84-87."""
    import torch.nn.functional as F
    return F.grid_sample(texture, sample_from_texture_dense(fim, wim, faces_uv_coord), align_corners=True)
