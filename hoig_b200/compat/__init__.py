"""Call-compatible stand-ins for the reference's three pybind extension modules (boundary B2).

``install()`` registers them in ``sys.modules`` under the names the reference Python imports
(``block_extractor_cuda``, ``local_attn_reshape_cuda``, ``neural_renderer.cuda.rasterize``), so
unmodified reference wrappers (thirdparty/block_extractor/block_extractor.py:1-54 etc.) run on
the B200 kernels.  Same ownership rules as the originals: the caller allocates and pre-fills
outputs, the ops write in place and return them; nothing is allocated here.
"""
from __future__ import annotations

import sys
import types

import torch

from .. import _lib, ops


def _stream():
    return torch.cuda.current_stream().cuda_stream


class block_extractor_cuda:  # noqa: N801 - mirrors the extension module name
    """thirdparty/block_extractor/block_extractor_cuda.cc:5-33."""

    @staticmethod
    def forward(source, flow_field, output, kernel_size):
        ops.block_extract(source, flow_field, output, kernel_size)
        return 1

    @staticmethod
    def backward(source, flow_field, grad_output, grad_source, grad_flow_field, kernel_size):
        ops.block_extract_backward(source, flow_field, grad_output, grad_source, grad_flow_field, kernel_size)
        return 1


class local_attn_reshape_cuda:  # noqa: N801
    """thirdparty/local_attn_reshape/local_attn_reshape_cuda.cc:5-29."""

    @staticmethod
    def forward(inputs, output, kernel_size):
        ops.local_attn_reshape(inputs, output, kernel_size)
        return 1

    @staticmethod
    def backward(inputs, grad_output, grad_inputs, kernel_size):
        ops.local_attn_reshape_backward(grad_output, grad_inputs, kernel_size)
        return 1


class rasterize:  # noqa: N801
    """thirdparty/neural_renderer/neural_renderer/cuda/rasterize_cuda.cpp:70-95,194-200."""

    @staticmethod
    def forward_face_index_map(faces, face_index_map, weight_map, depth_map, face_inv_map, faces_inv, image_size,
                               near, far, return_rgb, return_alpha, return_depth):
        for t, n in ((faces, "faces"), (face_index_map, "face_index_map"), (weight_map, "weight_map"),
                     (depth_map, "depth_map"), (faces_inv, "faces_inv")):
            if not t.is_cuda:
                raise RuntimeError(f"{n} must be a CUDA tensor")      # CHECK_CUDA, rasterize_cuda.cpp:66
            if not t.is_contiguous():
                raise RuntimeError(f"{n} must be contiguous")         # CHECK_CONTIGUOUS, rasterize_cuda.cpp:67
        if return_rgb or return_depth:
            raise NotImplementedError("hoig_b200: textured / depth-gradient rendering is outside the HOGAN hot path")
        B, F = faces.shape[:2]
        L = _lib.lib()
        _lib.check(L.hoig_face_inv(faces.data_ptr(), B * F, image_size, faces_inv.data_ptr(), _stream()), "face_inv")
        _lib.check(L.hoig_rasterize_fim_wim(faces.data_ptr(), B, F, image_size, near, far, 0, face_index_map.data_ptr(),
                                            weight_map.data_ptr(), depth_map.data_ptr(), None, 0, _stream()),
                   "rasterize_fim_wim")
        return [face_index_map, weight_map, depth_map, face_inv_map]


def _module(name, cls, fns):
    m = types.ModuleType(name)
    for f in fns:
        setattr(m, f, getattr(cls, f))
    return m


def install():
    """Register the stand-ins under the reference's extension-module names."""
    sys.modules["block_extractor_cuda"] = _module("block_extractor_cuda", block_extractor_cuda, ["forward", "backward"])
    sys.modules["local_attn_reshape_cuda"] = _module("local_attn_reshape_cuda", local_attn_reshape_cuda, ["forward", "backward"])
    sys.modules["neural_renderer.cuda.rasterize"] = _module("neural_renderer.cuda.rasterize", rasterize, ["forward_face_index_map"])
    return sys.modules["block_extractor_cuda"], sys.modules["local_attn_reshape_cuda"], sys.modules["neural_renderer.cuda.rasterize"]
