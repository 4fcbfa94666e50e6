"""Tensor-level wrappers over the C ABI (``include/hoig_b200.h``).

PyTorch is used for device memory and streams only.  Activations inside the
generator are NHWC tensors ``(N, H, W, C)`` whose channel dimension may be a
slice of a wider buffer (``stride(2) = ld >= C``).  Every function launches on
the current torch stream and raises ``RuntimeError`` on a non-zero status.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, CONV, CONV_LOCAL_ATTN,  # noqa: F401
                   CONV_TRANSPOSED, HOIG_BF16, HOIG_F16, HOIG_F32, ConvDesc)

_DT = {torch.float32: HOIG_F32, torch.bfloat16: HOIG_BF16, torch.float16: HOIG_F16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"hoig_b200: unsupported dtype {t.dtype}") from None


def _nhwc(t: torch.Tensor, name: str):
    """(ptr, ld) of an NHWC view; validates the layout."""
    if t.dim() != 4 or not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA (N,H,W,C) tensor, got {tuple(t.shape)} on {t.device}")
    n, h, w, c = t.shape
    ld = t.stride(2)
    if t.stride(3) != 1 or t.stride(1) != w * ld or (n > 1 and t.stride(0) != h * w * ld) or ld < c:
        raise ValueError(f"{name}: not an NHWC view (shape {tuple(t.shape)}, strides {t.stride()})")
    return t.data_ptr(), ld


def _f32c(t: torch.Tensor, name: str) -> int:
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise ValueError(f"{name}: expected a contiguous CUDA float32 tensor")
    return t.data_ptr()


_PACKED_DIMS: dict = {}


def packed_dims(cout: int, kh: int, kw: int, cin: int, mode: int = CONV, stride: int = 1, pad: int = 0):
    key = (cout, kh, kw, cin, mode, stride, pad)
    hit = _PACKED_DIMS.get(key)                # a pure function of the geometry: one C call per distinct conv
    if hit is None:
        r, c = ctypes.c_int(), ctypes.c_int()
        _lib.load().hoig_conv_packed_dims(mode, cout, kh, kw, cin, stride, pad, ctypes.byref(r), ctypes.byref(c))
        hit = _PACKED_DIMS[key] = (r.value, c.value)
    return hit


# ------------------------------------------------------------------ stage R
def rasterize(faces: torch.Tensor, image_size: int = 256, near: float = 0.1, far: float = 100.0,
              flip_y: bool = True, return_depth: bool = False, use_workspace: bool = True):
    """faces (B,F,3,3) f32 -> fim (B,is,is) int32, wim (B,is,is,3) f32[, depth (B,is,is)]."""
    B, F = faces.shape[:2]
    fp = _f32c(faces, "faces")
    fim = torch.empty(B, image_size, image_size, dtype=torch.int32, device=faces.device)
    wim = torch.empty(B, image_size, image_size, 3, dtype=torch.float32, device=faces.device)
    depth = torch.empty(B, image_size, image_size, dtype=torch.float32, device=faces.device) if return_depth else None
    L = _lib.lib()
    ws_bytes = L.hoig_rasterize_workspace_bytes(B, F, image_size) if use_workspace else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=faces.device) if ws_bytes else None
    _lib.check(L.hoig_rasterize_fim_wim(fp, B, F, image_size, near, far, int(flip_y), fim.data_ptr(), wim.data_ptr(),
                                        depth.data_ptr() if depth is not None else None,
                                        ws.data_ptr() if ws is not None else None, ws_bytes, _stream()),
               "rasterize_fim_wim")
    return (fim, wim, depth) if return_depth else (fim, wim)


def face_inv(faces: torch.Tensor, image_size: int) -> torch.Tensor:
    out = torch.zeros(faces.shape[0], faces.shape[1], 9, dtype=torch.float32, device=faces.device)
    _lib.check(_lib.lib().hoig_face_inv(_f32c(faces, "faces"), faces.shape[0] * faces.shape[1], image_size,
                                        out.data_ptr(), _stream()), "face_inv")
    return out


def project_faces(verts: torch.Tensor, cam: torch.Tensor, faces_idx: torch.Tensor, eye_z: float) -> torch.Tensor:
    B, V = verts.shape[:2]
    F = faces_idx.shape[0]
    assert faces_idx.dtype == torch.int32 and faces_idx.is_contiguous()
    out = torch.empty(B, F, 3, 3, dtype=torch.float32, device=verts.device)
    _lib.check(_lib.lib().hoig_project_faces(_f32c(verts, "verts"), _f32c(cam, "cam"), faces_idx.data_ptr(), B, V, F,
                                             eye_z, out.data_ptr(), _stream()), "project_faces")
    return out


def condition_maps(fim: torch.Tensor, map_fn: torch.Tensor, sem_full: torch.Tensor, n_hand_faces: int):
    B, H, W = fim.shape
    F = map_fn.shape[0] - 1
    dev = fim.device
    cond = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
    seg = torch.empty(B, 15, H, W, dtype=torch.float32, device=dev)
    not_hand = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().hoig_condition_maps(fim.data_ptr(), B, F, H, _f32c(map_fn, "map_fn"),
                                              _f32c(sem_full.reshape(-1), "sem_full"), n_hand_faces, cond.data_ptr(),
                                              seg.data_ptr(), not_hand.data_ptr(), _stream()), "condition_maps")
    return cond, seg, not_hand


def bc_transform(src_faces: torch.Tensor, fim_ref: torch.Tensor, wim_ref: torch.Tensor) -> torch.Tensor:
    B, F = src_faces.shape[:2]
    H = fim_ref.shape[1]
    T = torch.empty(B, H, H, 2, dtype=torch.float32, device=fim_ref.device)
    _lib.check(_lib.lib().hoig_bc_transform(_f32c(src_faces, "src_faces"), fim_ref.data_ptr(), _f32c(wim_ref, "wim"),
                                            B, F, H, T.data_ptr(), _stream()), "bc_transform")
    return T


def erode(mask: torch.Tensor, ks: int) -> torch.Tensor:
    B, _, H, W = mask.shape
    out = torch.empty_like(mask)
    _lib.check(_lib.lib().hoig_erode(_f32c(mask, "mask"), out.data_ptr(), B, H, W, ks, _stream()), "erode")
    return out


def condition_inputs(src_img, src_faces, fim_src, fim_ref, wim_ref, map_fn, sem_full, render_src, render_ref,
                     n_hand_faces: int, bg_erode_ks: int = 15):
    """``hoig_condition_inputs``: every Generator.forward input and the four masks in one launch (models/trainer.py:66-145)."""
    B, H, _ = fim_src.shape
    F = map_fn.shape[0] - 1
    dev = fim_src.device
    for t in (fim_src, fim_ref):
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise ValueError("condition_inputs: face-index maps must be contiguous int32")

    def new(c):
        return torch.empty(B, c, H, H, dtype=torch.float32, device=dev)

    out = dict(bg_inputs=new(4), src_obj_inputs=new(3), src_obj_conds=new(12), src_hand_inputs=new(3), src_hand_conds=new(3),
               tsf_obj_inputs=new(3), tsf_obj_conds=new(12), tsf_hand_inputs=new(3), tsf_hand_conds=new(3),
               T=torch.empty(B, H, H, 2, dtype=torch.float32, device=dev))
    masks = dict(src_mask_bg=new(1), ref_mask_bg=new(1), src_mask_hand=new(1), ref_mask_hand=new(1))
    d = _lib.CondInputsDesc()
    d.fim_src, d.fim_ref = fim_src.data_ptr(), fim_ref.data_ptr()
    d.wim_ref, d.src_faces, d.src_img = _f32c(wim_ref, "wim_ref"), _f32c(src_faces, "src_faces"), _f32c(src_img, "src_img")
    d.render_src, d.render_ref = _f32c(render_src, "render_src"), _f32c(render_ref, "render_ref")
    d.map_fn, d.sem_full = _f32c(map_fn, "map_fn"), _f32c(sem_full.reshape(-1), "sem_full")
    for k, v in {**out, **masks}.items():
        setattr(d, k, v.data_ptr())
    d.B, d.F, d.image_size, d.n_hand_faces, d.bg_erode_ks = B, F, H, n_hand_faces, bg_erode_ks
    _lib.check(_lib.lib().hoig_condition_inputs(ctypes.byref(d), _stream()), "condition_inputs")
    return out, masks


def uv_backward_warp(src_faces: torch.Tensor, fim_uv: torch.Tensor, wim_uv: torch.Tensor, src_fim: torch.Tensor):
    """utils/nmr.py:973-1040: (T (B,Hu,Wu,2), O (B,1,Hu,Wu)) of the atlas pixels, see ``hoig_uv_backward_warp``."""
    B, F = src_faces.shape[:2]
    Hu, Wu = fim_uv.shape[-2:]
    if fim_uv.dtype != torch.int32 or src_fim.dtype != torch.int32 or not fim_uv.is_contiguous() or not src_fim.is_contiguous():
        raise ValueError("uv_backward_warp: face-index maps must be contiguous int32")
    T = torch.empty(B, Hu, Wu, 2, dtype=torch.float32, device=src_faces.device)
    O = torch.empty(B, 1, Hu, Wu, dtype=torch.float32, device=src_faces.device)
    _lib.check(_lib.lib().hoig_uv_backward_warp(_f32c(src_faces, "src_faces"), fim_uv.data_ptr(), _f32c(wim_uv, "wim_uv"),
                                                src_fim.data_ptr(), B, F, Hu, Wu, src_fim.shape[-1], T.data_ptr(), O.data_ptr(),
                                                _stream()), "uv_backward_warp")
    return T, O


def sample_texture_dense(uv_coord: torch.Tensor, fim: torch.Tensor, wim: torch.Tensor) -> torch.Tensor:
    """utils/nmr.py:1068-1100: T (B,H,W,2) = barycentric blend of the per-face UV coordinates, -2 off the mesh."""
    B, H, W = fim.shape
    if fim.dtype != torch.int32 or not fim.is_contiguous():
        raise ValueError("sample_texture_dense: fim must be contiguous int32")
    T = torch.empty(B, H, W, 2, dtype=torch.float32, device=fim.device)
    _lib.check(_lib.lib().hoig_sample_texture_dense(_f32c(uv_coord, "uv_coord"), fim.data_ptr(), _f32c(wim, "wim"), B, H, W,
                                                    T.data_ptr(), _stream()), "sample_texture_dense")
    return T


def grid_sample_nchw(im: torch.Tensor, grid: torch.Tensor, align_corners: bool) -> torch.Tensor:
    """F.grid_sample(im, grid, 'bilinear', 'zeros', align_corners) for NCHW f32 images."""
    B, C, Hi, Wi = im.shape
    _, Ho, Wo, two = grid.shape
    if two != 2 or grid.shape[0] != B:
        raise ValueError("grid_sample_nchw: grid must be (B,Ho,Wo,2)")
    out = torch.empty(B, C, Ho, Wo, dtype=torch.float32, device=im.device)
    _lib.check(_lib.lib().hoig_grid_sample_nchw(_f32c(im, "im"), B, C, Hi, Wi, _f32c(grid, "grid"), Ho, Wo, int(align_corners),
                                                out.data_ptr(), _stream()), "grid_sample_nchw")
    return out


def uv_texture_compose(syn: torch.Tensor, O: torch.Tensor, preload: Optional[torch.Tensor] = None, x0: int = 384) -> torch.Tensor:
    """utils/nmr.py:1049-1056 in place on ``syn`` (B,C,Hu,Wu): open3(O) blend towards 1, then the stock object texture."""
    B, C, Hu, Wu = syn.shape
    if preload is not None and tuple(preload.shape) != (Hu, Wu - x0, C):
        raise ValueError(f"uv_texture_compose: preload must be ({Hu},{Wu - x0},{C}) HWC")
    _lib.check(_lib.lib().hoig_uv_texture_compose(_f32c(syn, "syn"), _f32c(O, "O"), _f32c(preload, "preload") if preload is not None else None,
                                                  B, C, Hu, Wu, x0 if preload is not None else Wu, _stream()), "uv_texture_compose")
    return syn


# ------------------------------------------------------- reference op boundary
def block_extract(source: torch.Tensor, flow: torch.Tensor, out: torch.Tensor, k: int) -> torch.Tensor:
    B, C, Hs, Ws = source.shape
    _, two, Hf, Wf = flow.shape
    if two != 2 or out.shape != (B, C, k * Hf, k * Wf):
        raise ValueError("block_extract: shape mismatch")
    _lib.check(_lib.lib().hoig_block_extract_f32(_f32c(source, "source"), _f32c(flow, "flow"), _f32c(out, "output"),
                                                 B, C, Hs, Ws, Hf, Wf, k, _stream()), "block_extract_f32")
    return out


def block_extract_backward(source, flow, grad_out, grad_source, grad_flow, k: int):
    """Adds the BlockExtractor gradients into ``grad_source`` / ``grad_flow`` (block_extractor_kernel.cu:86-166)."""
    B, C, Hs, Ws = source.shape
    _, two, Hf, Wf = flow.shape
    if two != 2 or grad_out.shape != (B, C, k * Hf, k * Wf) or grad_source.shape != source.shape or grad_flow.shape != flow.shape:
        raise ValueError("block_extract_backward: shape mismatch")
    _lib.check(_lib.lib().hoig_block_extract_backward_f32(_f32c(source, "source"), _f32c(flow, "flow"), _f32c(grad_out, "grad_output"),
                                                          _f32c(grad_source, "grad_source"), _f32c(grad_flow, "grad_flow_field"),
                                                          B, C, Hs, Ws, Hf, Wf, k, _stream()), "block_extract_backward_f32")
    return grad_source, grad_flow


def local_attn_reshape_backward(grad_out: torch.Tensor, grad_in: torch.Tensor, k: int) -> torch.Tensor:
    B, kk, H, W = grad_in.shape
    if kk != k * k or grad_out.shape != (B, 1, k * H, k * W):
        raise ValueError("local_attn_reshape_backward: shape mismatch")
    _lib.check(_lib.lib().hoig_local_attn_reshape_backward_f32(_f32c(grad_out, "grad_output"), _f32c(grad_in, "grad_inputs"), B, k, H, W,
                                                               _stream()), "local_attn_reshape_backward_f32")
    return grad_in


def local_attn_reshape(inputs: torch.Tensor, out: torch.Tensor, k: int) -> torch.Tensor:
    B, kk, H, W = inputs.shape
    if kk != k * k or out.shape != (B, 1, k * H, k * W):
        raise ValueError("local_attn_reshape: shape mismatch")
    _lib.check(_lib.lib().hoig_local_attn_reshape_f32(_f32c(inputs, "inputs"), _f32c(out, "output"), B, k, H, W,
                                                      _stream()), "local_attn_reshape_f32")
    return out


# ------------------------------------------------------------------ stage G
def conv2d(x0: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, *, kh: int, kw: int, stride: int = 1,
           pad: int = 0, pad_w: Optional[int] = None, mode: int = CONV, x1: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
           act: int = ACT_NONE, residual: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
           flow: Optional[torch.Tensor] = None, cout: Optional[int] = None, simt: bool = False,
           act_table: Optional[torch.Tensor] = None, spade_x: Optional[torch.Tensor] = None,
           spade_stats: Optional[torch.Tensor] = None, eps: float = 1e-5) -> torch.Tensor:
    """Implicit-GEMM convolution; see ``hoigConvDesc``.  With ``spade_x`` the GEMM produces interleaved (gamma, beta) and the
    epilogue writes relu(norm(spade_x) * (1 + gamma) + beta) into ``out`` (cout // 2 channels; pass ``cout``).  ``weight`` is the packed matrix
    from :func:`hoig_b200.packing.pack_conv_weight`; ``out`` is an NHWC view."""
    d = ConvDesc()
    d.dtype, d.mode = _dt(x0), mode
    N, H, W, C0 = x0.shape
    d.N, d.H, d.W, d.C0 = N, H, W, C0
    d.src0, d.ld0 = _nhwc(x0, "x0")
    if x1 is not None:
        if x1.shape[:3] != x0.shape[:3] or x1.dtype != x0.dtype:
            raise ValueError("conv2d: x1 must match x0 in batch/spatial size and dtype")
        d.C1 = x1.shape[3]
        d.src1, d.ld1 = _nhwc(x1, "x1")
    else:
        d.C1, d.src1, d.ld1 = 0, None, 0
    No, OH, OW, Co = out.shape
    d.OH, d.OW, d.Cout = OH, OW, (cout if cout is not None else Co)
    d.KH, d.KW, d.stride, d.pad = kh, kw, stride, pad
    d.pad_w = pad if pad_w is None else pad_w
    rows, cols = packed_dims(d.Cout, kh, kw, d.C0 + d.C1, mode, stride, pad)
    if tuple(weight.shape) != (rows, cols) or weight.dtype != x0.dtype or not weight.is_contiguous():
        raise ValueError(f"conv2d: packed weight must be {(rows, cols)} {x0.dtype}, got {tuple(weight.shape)} {weight.dtype}")
    if No != N or out.dtype != x0.dtype:
        raise ValueError("conv2d: output batch/dtype mismatch")
    d.weight = weight.data_ptr()
    d.bias = _f32c(bias, "bias") if bias is not None else None
    d.act = act
    if residual is not None:
        d.residual, d.ldr = _nhwc(residual, "residual")
    else:
        d.residual, d.ldr = None, 0
    d.dst, d.ldd = _nhwc(out, "out")
    if spade_x is not None:
        if spade_stats is None or spade_stats.dtype != torch.float64 or spade_stats.numel() != N * (d.Cout // 2) * 2:
            raise ValueError("conv2d: spade_stats must be float64 [N, Cout/2, 2]")
        if spade_x.shape != (N, OH, OW, d.Cout // 2) or spade_x.dtype != x0.dtype or Co < d.Cout // 2:
            raise ValueError("conv2d: spade_x / out must be (N,OH,OW,Cout/2) of the input dtype")
        d.spade_x, d.ld_spade_x = _nhwc(spade_x, "spade_x")
        d.spade_stats, d.spade_eps = spade_stats.data_ptr(), eps
    else:
        d.spade_x, d.ld_spade_x, d.spade_stats, d.spade_eps = None, 0, None, 0.0
    d.stats = stats.data_ptr() if stats is not None else None
    if stats is not None and (stats.dtype != torch.float64 or stats.numel() != N * d.Cout * 2):
        raise ValueError("conv2d: stats must be float64 [N, Cout, 2]")
    d.flow = _f32c(flow, "flow") if flow is not None else None
    if act_table is not None:
        if act_table.dtype != torch.int32 or act_table.numel() != d.Cout or not act_table.is_cuda:
            raise ValueError("conv2d: act_table must be a CUDA int32 tensor with Cout entries")
        d.act_table = act_table.data_ptr()
    else:
        d.act_table = None
    L = _lib.lib()
    if _lib.recorder.timing:
        _lib.recorder.tag = (f"{('conv', 'convT', 'attn')[mode]} k{kh}x{kw} s{stride} Cin{d.C0 + d.C1} Cout{d.Cout} "
                             f"{H}x{W}->{OH}x{OW} N{N}" + (" spade" if spade_x is not None else ""))
    fn = L.hoig_conv2d_simt if simt else L.hoig_conv2d
    _lib.check(fn(ctypes.byref(d), _stream()), "conv2d")
    return out


def nchw_to_nhwc(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    B, C, H, W = x.shape
    ptr, ld = _nhwc(out, "out")
    _lib.check(_lib.lib().hoig_nchw_to_nhwc(_f32c(x, "x"), B, C, H, W, ptr, ld, out.shape[3], _dt(out), _stream()),
               "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x: torch.Tensor, channels: int) -> torch.Tensor:
    N, H, W, _ = x.shape
    ptr, ld = _nhwc(x, "x")
    out = torch.empty(N, channels, H, W, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().hoig_nhwc_to_nchw(ptr, ld, _dt(x), N, channels, H, W, out.data_ptr(), _stream()), "nhwc_to_nchw")
    return out


def seg_resize(seg: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    B, C, Hi, Wi = seg.shape
    _, Ho, Wo, Cpad = out.shape
    ptr, ld = _nhwc(out, "out")
    _lib.check(_lib.lib().hoig_seg_resize_nearest(_f32c(seg, "seg"), B, C, Hi, Wi, ptr, ld, Cpad, Ho, Wo, _dt(out),
                                                  _stream()), "seg_resize_nearest")
    return out


def seg_unfold3(seg: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """seg NCHW f32 -> out (B,Ho,Wo,Kpad): nearest resize + 3x3 im2col, channel t*C + c (zero padded), see hoig_b200.h."""
    B, C, Hi, Wi = seg.shape
    _, Ho, Wo, Kpad = out.shape
    ptr, ld = _nhwc(out, "out")
    _lib.check(_lib.lib().hoig_seg_unfold3(_f32c(seg, "seg"), B, C, Hi, Wi, ptr, ld, Kpad, Ho, Wo, _dt(out), _stream()),
               "seg_unfold3")
    return out


def plane_stats(x: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    N, H, W, C = x.shape
    ptr, ld = _nhwc(x, "x")
    _lib.check(_lib.lib().hoig_plane_stats(ptr, ld, _dt(x), N, H * W, C, stats.data_ptr(), _stream()), "plane_stats")
    return stats


def instnorm_apply(x: torch.Tensor, stats: torch.Tensor, out: torch.Tensor, *, gamma=None, beta=None, gb=None,
                   residual=None, relu: bool = False, eps: float = 1e-5) -> torch.Tensor:
    N, H, W, C = x.shape
    xp, ldx = _nhwc(x, "x")
    op, ldo = _nhwc(out, "out")
    gp, ldg = _nhwc(gb, "gb") if gb is not None else (None, 0)
    rp, ldr = _nhwc(residual, "residual") if residual is not None else (None, 0)
    if _lib.recorder.timing:
        _lib.recorder.tag = f"C{C} {H}x{W} N{N} gb={int(gb is not None)} res={int(residual is not None)} ld={ldx}->{ldo}"
    _lib.check(_lib.lib().hoig_instnorm_apply(xp, ldx, stats.data_ptr(),
                                              _f32c(gamma, "gamma") if gamma is not None else None,
                                              _f32c(beta, "beta") if beta is not None else None,
                                              gp, ldg, rp, ldr, int(relu), op, ldo, _dt(x), N, H * W, C, eps, _stream()),
               "instnorm_apply")
    return out


def resize_flow(T: torch.Tensor, h: int, subtract_identity: bool = True) -> torch.Tensor:
    B, Hi, Wi, _ = T.shape
    flow = torch.empty(B, h, h, 2, dtype=torch.float32, device=T.device)
    _lib.check(_lib.lib().hoig_resize_flow(_f32c(T, "T"), B, Hi, Wi, h, int(subtract_identity), flow.data_ptr(),
                                           _stream()), "resize_flow")
    return flow


def attn_finish(hidden: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, src: torch.Tensor, flow: torch.Tensor,
                tgt: torch.Tensor, out: torch.Tensor, k: int, unfold: Optional[torch.Tensor] = None) -> torch.Tensor:
    N, h, _, C = src.shape
    hp, ldh = _nhwc(hidden, "hidden")
    sp, lds = _nhwc(src, "src")
    tp, ldt = _nhwc(tgt, "tgt")
    op, ldo = _nhwc(out, "out")
    up, ldu = _nhwc(unfold, "unfold") if unfold is not None else (None, 0)
    _lib.check(_lib.lib().hoig_attn_finish(hp, ldh, hidden.shape[3], _f32c(w2, "w2"), _f32c(b2, "b2"), sp, lds,
                                           _f32c(flow, "flow"), tp, ldt, op, ldo, _dt(src), N, h, C, k, up, ldu, _stream()),
               "attn_finish")
    return out


def attn_unfold(src: torch.Tensor, tgt: torch.Tensor, flow: torch.Tensor, out: torch.Tensor, k: int) -> torch.Tensor:
    """out (N,h,h,2*k*k*C): per tap t the channels [BlockExtractor(tgt,0) tap | BlockExtractor(src,flow) tap]."""
    N, h, _, C = src.shape
    sp, lds = _nhwc(src, "src")
    tp, ldt = _nhwc(tgt, "tgt")
    op, ldo = _nhwc(out, "out")
    _lib.check(_lib.lib().hoig_attn_unfold(sp, lds, tp, ldt, _f32c(flow, "flow"), op, ldo, _dt(src), N, h, C, k, _stream()),
               "attn_unfold")
    return out


def replicate_pad(x: torch.Tensor, out: torch.Tensor, pad: int) -> torch.Tensor:
    """out (N,h+2p,h+2p,C) = x with its edge pixels replicated ``pad`` times (BlockExtractor's clamped taps)."""
    N, h, _, C = x.shape
    xp, ldx = _nhwc(x, "x")
    op, ldo = _nhwc(out, "out")
    if out.shape != (N, h + 2 * pad, h + 2 * pad, C):
        raise ValueError(f"replicate_pad: out has shape {tuple(out.shape)}")
    _lib.check(_lib.lib().hoig_replicate_pad(xp, ldx, op, ldo, _dt(x), N, h, C, pad, _stream()), "replicate_pad")
    return out


def conv2d_halo(segments, kh: int, kw: int, cout: int):
    """Dense kh x kw stride-1 conv over padded rasters on the tensor cores (raw accumulators, 16-bit dtypes).

    ``segments``: one or two ``(x, packed_weight, out)`` with x (N,Hp,Wp,C) and out (N,Hp,Wp,cout); both run in ONE launch.
    Interior pixels (>= k//2 from the raster border) get the exact convolution, border pixels are don't-care."""
    segs = (_lib.HaloConvSeg * len(segments))()
    for sg, (x, w, out) in zip(segs, segments):
        xp, ldx = _nhwc(x, "x")
        op, ldo = _nhwc(out, "out")
        N, Hp, Wp, C = x.shape
        if out.shape[:3] != x.shape[:3] or out.shape[3] != cout or w.dtype != x.dtype or out.dtype != x.dtype:
            raise ValueError("conv2d_halo: out must be (N,Hp,Wp,cout) of the input dtype")
        if tuple(w.shape) != (cout, kh * kw * C) or not w.is_contiguous():
            raise ValueError(f"conv2d_halo: packed weight must be ({cout},{kh * kw * C}), got {tuple(w.shape)}")
        sg.src, sg.ld, sg.N, sg.Hp, sg.Wp, sg.C = xp, ldx, N, Hp, Wp, C
        sg.weight, sg.dst, sg.ldd = w.data_ptr(), op, ldo
    dt = _dt(segments[0][0])
    _lib.check(_lib.lib().hoig_conv2d_halo(dt, kh, kw, cout, segs, len(segments), _stream()), "conv2d_halo")
    return [s[2] for s in segments]


def attn_combine(gt: torch.Tensor, gs: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, src: torch.Tensor,
                 flow: torch.Tensor, tgt: torch.Tensor, out: torch.Tensor, k: int) -> torch.Tensor:
    """out = tgt + local-attention warp of src, from the commuted k x k convs gt (target) and gs (source, extended grid)."""
    N, h, _, C = src.shape
    gtp, ldgt = _nhwc(gt, "gt")
    gsp, ldgs = _nhwc(gs, "gs")
    sp, lds = _nhwc(src, "src")
    tp, ldt = _nhwc(tgt, "tgt")
    op, ldo = _nhwc(out, "out")
    if gt.shape[1] != h + k - 1 or gs.shape[1] != h + 2 * (k - 1) or gt.shape[3] != gs.shape[3]:
        raise ValueError("attn_combine: gt must be (N,h+k-1,h+k-1,Chid) and gs (N,h+2k-2,h+2k-2,Chid)")
    _lib.check(_lib.lib().hoig_attn_combine(gtp, ldgt, gsp, ldgs, gt.shape[3], _f32c(b1, "b1"), _f32c(w2, "w2"), _f32c(b2, "b2"),
                                            sp, lds, _f32c(flow, "flow"), tp, ldt, op, ldo, _dt(src), N, h, C, k, _stream()),
               "attn_combine")
    return out


def grid_sample(x: torch.Tensor, grid: torch.Tensor, out: torch.Tensor, tgt: Optional[torch.Tensor] = None):
    N, h, _, C = x.shape
    xp, ldx = _nhwc(x, "x")
    op, ldo = _nhwc(out, "out")
    tp, ldt = _nhwc(tgt, "tgt") if tgt is not None else (None, 0)
    _lib.check(_lib.lib().hoig_grid_sample(xp, ldx, _f32c(grid, "grid"), tp, ldt, op, ldo, _dt(x), N, h, C, _stream()),
               "grid_sample")
    return out


def hunfold_nchw(x: torch.Tensor, out: torch.Tensor, k: int) -> torch.Tensor:
    """NCHW f32 (B,C,H,W) -> NHWC (B,H,W,Cpad) with channel s*C + c = x[b,c,y,x+s-k//2] (zero padded)."""
    B, C, H, W = x.shape
    ptr, ld = _nhwc(out, "out")
    _lib.check(_lib.lib().hoig_hunfold_nchw(_f32c(x, "x"), B, C, H, W, k, ptr, ld, out.shape[3], _dt(out), _stream()),
               "hunfold_nchw")
    return out


def hfold_nchw(z: torch.Tensor, groups: int, k: int, segments, act_table: Optional[torch.Tensor] = None):
    """Z (B,H,W,>=k*G) -> list of NCHW f32 tensors, one per (c0, n) segment of the G folded channels."""
    B, H, W, _ = z.shape
    zp, ldz = _nhwc(z, "z")
    outs = [torch.empty(B, n, H, W, dtype=torch.float32, device=z.device) for _, n in segments]
    nseg = len(segments)
    arr_p = (ctypes.c_void_p * nseg)(*[o.data_ptr() for o in outs])
    arr_c0 = (ctypes.c_int * nseg)(*[c0 for c0, _ in segments])
    arr_n = (ctypes.c_int * nseg)(*[n for _, n in segments])
    _lib.check(_lib.lib().hoig_hfold_nchw(zp, ldz, _dt(z), B, H, W, groups, k,
                                          act_table.data_ptr() if act_table is not None else None, nseg, arr_p, arr_c0,
                                          arr_n, _stream()), "hfold_nchw")
    return outs


def composite(img_bg, obj, hand, mask_bg, mask_hand) -> torch.Tensor:
    B, _, H, W = img_bg.shape
    out = torch.empty_like(img_bg)
    _lib.check(_lib.lib().hoig_composite(_f32c(img_bg, "img_bg"), _f32c(obj, "obj"), _f32c(hand, "hand"),
                                         _f32c(mask_bg, "mask_bg"), _f32c(mask_hand, "mask_hand"), out.data_ptr(), B,
                                         H * W, _stream()), "composite")
    return out
