"""Stage R host side: batched replacements for the reference's per-sample renderer calls.

Boundary B3 of SURVEY.md section 8b (paths relative to /root/reference/HOIG_HOv3):
  * ``nr.rasterize_face_index_map_and_weight_map`` (thirdparty/neural_renderer/neural_renderer/rasterize.py:543-571)
  * ``MANORenderer.render_fim_wim``                (utils/nmr.py:496-513)
  * ``MANORenderer.encode_fim / encode_sem``       (utils/nmr.py:567-595)
  * ``MANORenderer.cal_bc_transform``              (utils/nmr.py:874-968; T only, O is discarded by the caller)
  * ``util.morph(mode='erode')``                   (utils/util.py:142-153)
The reference loops over the batch in Python (models/trainer.py:63-97); here every call is one
launch over the whole batch.
"""
from __future__ import annotations

import math

import torch

from . import ops

EYE_Z = -(1.0 / math.tan(math.radians(30.0)) + 1.0)  # utils/nmr.py:357
N_HAND_FACES = 1538                                  # models/trainer.py:72


def rasterize_face_index_map_and_weight_map(faces, image_size=256, anti_aliasing=False, near=0.1, far=100.0, eps=1e-4):
    """Same result as the reference function: ``(fim (B,is,is) int32, wim (B,is,is,3) f32)``, already flipped."""
    if anti_aliasing:
        raise NotImplementedError("HOGAN calls the face-index rasterizer with anti_aliasing=False (utils/nmr.py:512)")
    return ops.rasterize(faces.contiguous().float(), image_size, near, far, flip_y=True)


def render_fim_wim_batched(cam: torch.Tensor, vertices: torch.Tensor, faces_idx: torch.Tensor, image_size: int = 256):
    """cam (B,15), vertices (B,V,3), faces_idx (F,3) int32 -> (faces (B,F,3,3), fim, wim)."""
    faces = ops.project_faces(vertices.contiguous().float(), cam.contiguous().float(), faces_idx.contiguous().int(), EYE_Z)
    fim, wim = ops.rasterize(faces, image_size)
    return faces, fim, wim


def condition_inputs_fused(src_img, faces_src, fim_src, fim_ref, wim_ref, map_fn, sem_full, render_img_src=None,
                           render_img_ref=None, n_hand_faces: int = N_HAND_FACES):
    """Same result as :func:`condition_inputs` from ONE kernel launch (``hoig_condition_inputs``, row N1 of SURVEY 8f)."""
    r_src = src_img if render_img_src is None else render_img_src
    r_ref = src_img if render_img_ref is None else render_img_ref
    return ops.condition_inputs(src_img.contiguous(), faces_src, fim_src, fim_ref, wim_ref, map_fn, sem_full, r_src.contiguous(),
                                r_ref.contiguous(), n_hand_faces)


def condition_inputs(src_img, faces_src, fim_src, fim_ref, wim_ref, map_fn, sem_full, render_img_src=None,
                     render_img_ref=None, n_hand_faces: int = N_HAND_FACES):
    """Batched ``HandRecoveryFlow.forward`` tail (models/trainer.py:66-145) given the rasterizer outputs.

    ``render_img_*`` are the UV-texture re-renderings (stage R8, a "next" row); when omitted the source image
    stands in so the tensor shapes match.  Returns the dict of generator inputs plus the masks.
    """
    out = {}
    side = {}
    for tag, fim in (("src", fim_src), ("ref", fim_ref)):
        cond, seg, not_hand = ops.condition_maps(fim, map_fn, sem_full, n_hand_faces)
        m_hand = ops.erode(not_hand, 3)                                   # trainer.py:72
        m_bg = ops.erode(cond[:, -1:].contiguous(), 3)                    # trainer.py:109-110
        hm = (cond[:, :1] < 1.5).float()                                  # trainer.py:112-124
        om = (cond[:, :1] > 1.5).float()
        cond_hand = torch.cat([hm * cond[:, :2], cond[:, 2:] + 1 - hm], 1)
        cond_obj = torch.cat([om * cond[:, :2], cond[:, 2:] + 1 - om], 1)
        side[tag] = dict(cond=cond, seg=seg, m_hand=m_hand, m_bg=m_bg, cond_hand=cond_hand, cond_obj=cond_obj)
    T = ops.bc_transform(faces_src, fim_ref, wim_ref)                     # nmr.py:874-925 (+ y flip of trainer.py:67-68)
    mh = side["ref"]["m_hand"][:, 0][:, :, :, None]
    T_hand = torch.where(mh == 1, torch.full_like(T, -2.0), T)            # trainer.py:81
    r_src = src_img if render_img_src is None else render_img_src
    r_ref = src_img if render_img_ref is None else render_img_ref
    s, r = side["src"], side["ref"]
    bg_mask = ops.erode(s["cond"][:, -1:].contiguous(), 15)               # trainer.py:135
    out["bg_inputs"] = torch.cat([src_img * bg_mask, bg_mask], 1)
    out["src_obj_inputs"] = r_src * (s["m_hand"] - s["m_bg"])             # trainer.py:127 (rgb part)
    out["src_obj_conds"] = torch.cat([s["cond_obj"], s["seg"][:, 6:]], 1)
    out["src_hand_inputs"] = src_img * (1 - s["m_hand"])                  # trainer.py:128
    out["src_hand_conds"] = s["cond_hand"]
    out["tsf_obj_inputs"] = r_ref * (r["m_hand"] - r["m_bg"])             # trainer.py:131
    out["tsf_obj_conds"] = torch.cat([r["cond_obj"], r["seg"][:, 6:]], 1)
    out["tsf_hand_inputs"] = r_ref * (1 - r["m_hand"])                    # trainer.py:132
    out["tsf_hand_conds"] = r["cond_hand"]
    out["T"] = T_hand
    masks = dict(src_mask_bg=s["m_bg"], ref_mask_bg=r["m_bg"], src_mask_hand=s["m_hand"], ref_mask_hand=r["m_hand"])
    return out, masks


# ------------------------------------------------------------------ stage R8: UV-texture warp
def texture_backward_warp(im, src_faces, src_fim, fim_uv, wim_uv, obj_tex=None, x0: int = 384):
    """``MANORenderer.get_texture_backward_warp`` (utils/nmr.py:973-1058), batched.

    im (B,3,256,256) source images; src_faces (B,F,3,3) projected source faces (``hoig_project_faces`` layout; the y flip of
    trainer.py:67-68 happens inside); src_fim (B,256,256) int32; fim_uv / wim_uv the object's UV atlas maps (Hu,Wu) / (Hu,Wu,3);
    obj_tex the stock object texture (Hu, Wu-x0, 3) or None (``pre_load=False``).  Returns the texture atlas (B,3,Hu,Wu)."""
    T, O = ops.uv_backward_warp(src_faces, fim_uv, wim_uv, src_fim)
    syn = ops.grid_sample_nchw(im, T, align_corners=False)               # nmr.py:1047 (torch default)
    return ops.uv_texture_compose(syn, O, obj_tex, x0)                    # nmr.py:1049-1056


def sample_from_texture_dense(fim, wim, faces_uv_coord):
    """``MANORenderer.sample_from_texture_dense`` (utils/nmr.py:1068-1100), batched: T (B,H,W,2)."""
    return ops.sample_texture_dense(faces_uv_coord, fim, wim)


def render_from_texture(texture, fim, wim, faces_uv_coord):
    """models/trainer.py:84-87: re-render the texture atlas at a pose given its face-index / weight maps."""
    return ops.grid_sample_nchw(texture, sample_from_texture_dense(fim, wim, faces_uv_coord), align_corners=True)


# ------------------------------------------------------------------ row N1: the batched HandRecoveryFlow
class HandRecoveryFlowB200(torch.nn.Module):
    """Batched replacement for ``HandRecoveryFlow.forward`` (models/trainer.py:46-145) downstream of the MANO layer.

    The reference runs, per sample and in Python, two rasterizations, ~100 small kernels of table gathers / masks / boolean
    scatter and the UV-texture warp; here the whole batch is ~10 launches.  MANO/smplx (``HandModelRecovery.get_details``) is out
    of scope, so ``forward`` takes what ``get_details`` returns -- vertices and the 15-value camera row -- instead of MANO
    parameters.  The per-object tables are the reference's own buffers (utils/nmr.py:286-401), passed in unchanged:

      faces_idx (F,3) int32 ``faces_<obj>``; map_fn (F+1,3) ``map_fn_<obj>``; sem_full (F+1,1) ``sem_full_<obj>``;
      fim_uv (Hu,Wu) int32 / wim_uv (Hu,Wu,3) ``fim_uv_<obj>[0]`` / ``wim_uv_<obj>[0]``; faces_uv_coord (F,3,2)
      ``faces_uv_coord_<obj>[0]``; obj_tex (Hu, Wu-384, 3) ``obj_tex_img_<obj>`` (or None for ``pre_load=False``).

    One object per batch (the reference looks the object up per sample; group samples by object upstream)."""

    def __init__(self, faces_idx, map_fn, sem_full, fim_uv, wim_uv, faces_uv_coord, obj_tex=None, image_size: int = 256,
                 n_hand_faces: int = N_HAND_FACES, tex_x0: int = 384):
        super().__init__()
        self.image_size, self.n_hand_faces, self.tex_x0 = image_size, n_hand_faces, tex_x0
        self.n_verts = int(faces_idx.max().item()) + 1                          # trainer.py:65 ``length``
        self.register_buffer("faces_idx", faces_idx.contiguous().int())
        self.register_buffer("map_fn", map_fn.contiguous().float())
        self.register_buffer("sem_full", sem_full.contiguous().float())
        self.register_buffer("fim_uv", fim_uv.contiguous().int())
        self.register_buffer("wim_uv", wim_uv.contiguous().float())
        self.register_buffer("faces_uv_coord", faces_uv_coord.contiguous().float())
        self.register_buffer("obj_tex", None if obj_tex is None else obj_tex.contiguous().float())

    @torch.no_grad()
    def forward(self, src_img, src_verts, ref_verts, src_cam, ref_cam=None, src_armask=None, tsf_armask=None):
        """src_img (B,3,H,W); *_verts (B,V,3) (rows beyond the object's vertex count are ignored, data/hov3_dataset.py:246-248);
        *_cam (B,15).  Returns ``(generator_kwargs, masks)``: the keyword arguments of ``Generator.forward`` as built at
        trainer.py:377-393 and the four crop masks of trainer.py:144-145."""
        ref_cam = src_cam if ref_cam is None else ref_cam
        with torch.cuda.device(src_img.device):
            nv = self.n_verts
            vs = src_verts if src_verts.shape[1] == nv else src_verts[:, :nv].contiguous()
            vr = ref_verts if ref_verts.shape[1] == nv else ref_verts[:, :nv].contiguous()
            fs, fim_s, wim_s = render_fim_wim_batched(src_cam, vs, self.faces_idx, self.image_size)
            _, fim_r, wim_r = render_fim_wim_batched(ref_cam, vr, self.faces_idx, self.image_size)
            src_img = src_img.contiguous().float()
            tex = texture_backward_warp(src_img, fs, fim_s, self.fim_uv, self.wim_uv, self.obj_tex, self.tex_x0)   # trainer.py:83
            r_ref = render_from_texture(tex, fim_r, wim_r, self.faces_uv_coord)                                    # trainer.py:84-85
            r_src = render_from_texture(tex, fim_s, wim_s, self.faces_uv_coord)                                    # trainer.py:86-87
            kwargs, masks = condition_inputs_fused(src_img, fs, fim_s, fim_r, wim_r, self.map_fn, self.sem_full, r_src, r_ref,
                                                   self.n_hand_faces)
            if src_armask is not None:
                kwargs["src_armask"] = src_armask
            if tsf_armask is not None:
                kwargs["tsf_armask"] = tsf_armask
            return kwargs, masks
