"""ctypes binding of libhoig_b200.so (the C ABI in include/hoig_b200.h).

The library is built in-tree by ``hoig_b200/build.py`` (nvcc, sm_100a).  There
is no fallback: if it is missing or the device is not a B200-class GPU the
first op raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libhoig_b200.so")

HOIG_F32, HOIG_BF16, HOIG_F16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
CONV, CONV_TRANSPOSED, CONV_LOCAL_ATTN = 0, 1, 2


class ConvDesc(Structure):
    """Mirror of ``hoigConvDesc``."""
    _fields_ = [
        ("dtype", c_int), ("mode", c_int),
        ("N", c_int), ("H", c_int), ("W", c_int),
        ("C0", c_int), ("C1", c_int),
        ("OH", c_int), ("OW", c_int), ("Cout", c_int),
        ("KH", c_int), ("KW", c_int), ("stride", c_int), ("pad", c_int),
        ("src0", c_void_p), ("ld0", c_int64),
        ("src1", c_void_p), ("ld1", c_int64),
        ("weight", c_void_p),
        ("bias", c_void_p),
        ("act", c_int),
        ("residual", c_void_p), ("ldr", c_int64),
        ("dst", c_void_p), ("ldd", c_int64),
        ("stats", c_void_p),
        ("flow", c_void_p),
        ("act_table", c_void_p),
        ("pad_w", c_int),
        ("spade_x", c_void_p), ("ld_spade_x", c_int64),
        ("spade_stats", c_void_p),
        ("spade_eps", ctypes.c_float),
    ]


class CondInputsDesc(Structure):
    """Mirror of ``hoigCondInputsDesc``."""
    _fields_ = ([(n, c_void_p) for n in ("fim_src", "fim_ref", "wim_ref", "src_faces", "src_img", "render_src", "render_ref", "map_fn",
                                         "sem_full", "bg_inputs", "src_obj_inputs", "src_obj_conds", "src_hand_inputs", "src_hand_conds",
                                         "tsf_obj_inputs", "tsf_obj_conds", "tsf_hand_inputs", "tsf_hand_conds", "T", "src_mask_bg",
                                         "ref_mask_bg", "src_mask_hand", "ref_mask_hand")]
                + [(n, c_int) for n in ("B", "F", "image_size", "n_hand_faces", "bg_erode_ks")])


class HaloConvSeg(Structure):
    """Mirror of ``hoigHaloConvSeg``."""
    _fields_ = [("src", c_void_p), ("ld", c_int64), ("N", c_int), ("Hp", c_int), ("Wp", c_int), ("C", c_int),
                ("weight", c_void_p), ("dst", c_void_p), ("ldd", c_int64)]


_SIGNATURES = {
    # name: (restype, argtypes)
    "hoig_version": (c_char_p, []),
    "hoig_last_error": (c_char_p, []),
    "hoig_check_device": (c_int, []),
    "hoig_rasterize_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "hoig_rasterize_fim_wim": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_size_t, c_void_p]),
    "hoig_face_inv": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "hoig_project_faces": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "hoig_condition_maps": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "hoig_bc_transform": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hoig_erode": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "hoig_block_extract_f32": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "hoig_local_attn_reshape_f32": (c_int, [c_void_p, c_void_p] + [c_int] * 4 + [c_void_p]),
    "hoig_conv_packed_dims": (c_int, [c_int] * 7 + [POINTER(c_int), POINTER(c_int)]),
    "hoig_conv2d": (c_int, [POINTER(ConvDesc), c_void_p]),
    "hoig_conv2d_simt": (c_int, [POINTER(ConvDesc), c_void_p]),
    "hoig_set_umma_gather_only": (None, [c_int]),
    "hoig_set_rasterizer_band_pixels": (None, [c_int]),
    "hoig_nchw_to_nhwc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "hoig_nhwc_to_nchw": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hoig_seg_resize_nearest": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_int,
                                        c_int, c_void_p]),
    "hoig_plane_stats": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hoig_instnorm_apply": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                    c_int64, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "hoig_resize_flow": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hoig_attn_finish": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_void_p]),
    "hoig_attn_unfold": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "hoig_block_extract_backward_f32": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p]),
    "hoig_local_attn_reshape_backward_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "hoig_conv2d_wgrad_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p] + [c_int] * 12 + [c_void_p]),
    "hoig_instnorm_backward_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                           c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "hoig_condition_inputs": (c_int, [POINTER(CondInputsDesc), c_void_p]),
    "hoig_uv_backward_warp": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hoig_sample_texture_dense": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hoig_grid_sample_nchw": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hoig_uv_texture_compose": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hoig_seg_unfold3": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "hoig_replicate_pad": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hoig_conv2d_halo": (c_int, [c_int, c_int, c_int, c_int, POINTER(HaloConvSeg), c_int, c_void_p]),
    "hoig_set_halo_variant": (None, [c_int]),
    "hoig_set_umma_pair_mode": (None, [c_int]),
    "hoig_set_umma_dual_mode": (None, [c_int]),
    "hoig_set_umma_bres_mode": (None, [c_int]),
    "hoig_set_umma_halo_mode": (None, [c_int]),
    "hoig_set_umma_vhalo_mode": (None, [c_int]),
    "hoig_set_attn_tc_mode": (None, [c_int]),
    "hoig_attn_combine": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hoig_grid_sample": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "hoig_composite": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p]),
    "hoig_hunfold_nchw": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "hoig_hfold_nchw": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                POINTER(c_void_p), POINTER(c_int), POINTER(c_int), c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_device_checked = False


def load() -> ctypes.CDLL:
    """dlopen the library and bind every entry point (no CUDA call is made)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -m hoig_b200.build` (nvcc, sm_100a). "
                "hoig_b200 has no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().hoig_last_error().decode(errors="replace")
        raise RuntimeError(f"hoig_b200.{what} failed (status {status}): {msg}")


class LaunchRecorder:
    """Counts kernel launches made through the C ABI and, when ``timing`` is on, brackets each
    with CUDA events on the launching (current torch) stream.  Used by bench.py / profiling."""

    def __init__(self):
        self.launches = 0
        self.timing = False
        self.records = []   # (entry point, start event, end event, tag)
        self.tag = None     # optional label set by the caller for the next launch

    def reset(self, timing: bool = False):
        self.launches, self.timing, self.records, self.tag = 0, timing, [], None

    def summary(self):
        """{entry point: (launch count, total ms)} -- call after torch.cuda.synchronize()."""
        out = {}
        for name, s, e, _ in self.records:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + s.elapsed_time(e))
        return out

    def by_tag(self):
        """{(entry point, tag): (launch count, total ms)}."""
        out = {}
        for name, s, e, tag in self.records:
            n, t = out.get((name, tag), (0, 0.0))
            out[(name, tag)] = (n + 1, t + s.elapsed_time(e))
        return out


recorder = LaunchRecorder()
_NO_LAUNCH = {"hoig_version", "hoig_last_error", "hoig_check_device", "hoig_conv_packed_dims",
              "hoig_rasterize_workspace_bytes", "hoig_set_umma_gather_only", "hoig_set_rasterizer_band_pixels",
              "hoig_set_halo_variant", "hoig_set_umma_pair_mode", "hoig_set_umma_dual_mode", "hoig_set_umma_bres_mode",
              "hoig_set_umma_halo_mode", "hoig_set_umma_vhalo_mode", "hoig_set_attn_tc_mode"}


class _Proxy:
    def __init__(self, L):
        self._L = L

    def __getattr__(self, name):
        fn = getattr(self._L, name)
        if name in _NO_LAUNCH:
            return fn

        def launch(*args):
            rec = recorder
            if rec.timing:
                import torch
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = fn(*args)
                e.record()
                rec.records.append((name, s, e, rec.tag))
                rec.tag = None
            else:
                r = fn(*args)
            rec.launches += 1
            return r

        setattr(self, name, launch)
        return launch


_proxy = None


def lib():
    """Library handle for compute calls; verifies once that the device is sm_100."""
    global _device_checked, _proxy
    if _proxy is None:
        _proxy = _Proxy(load())
    if not _device_checked:
        check(load().hoig_check_device(), "check_device")
        _device_checked = True
    return _proxy
