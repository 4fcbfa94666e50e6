"""CPU parity oracle for the HOGAN generator hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``hoig_b200/`` may import this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker / reported baseline.

* ``hoig_oracle.c``      C restatement of the reference's three native ops
                         (rasterizer kernels 1+2, BlockExtractor fwd,
                         LocalAttnReshape fwd).
* ``generator_ref.py``   functional PyTorch-CPU restatement of
                         ``Generator.forward`` driven by a reference-layout
                         ``state_dict``.
* ``geometry_ref.py``    restatement of the HandRecoveryFlow geometry glue.
* ``build_ref.py``       compiles the reference's OWN CUDA kernels (sources
                         read in place from /root/reference, 10-site API shim)
                         into ``oracle/_ref`` -- the same-toolchain GPU oracle.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    """Compile hoig_oracle.c with gcc (seconds).  Returns the .so path."""
    so = os.path.join(_HERE, "_build", "libhoig_oracle.so")
    src = os.path.join(_HERE, "hoig_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libhoig_oracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        L.oracle_face_inv.argtypes = [f32p, ctypes.c_long, ctypes.c_int, f32p]
        L.oracle_face_inv.restype = None
        L.oracle_rasterize.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                       i32p, f32p, ctypes.c_void_p]
        L.oracle_rasterize.restype = None
        L.oracle_block_extract.argtypes = [f32p, f32p, f32p] + [ctypes.c_int] * 7
        L.oracle_block_extract.restype = None
        L.oracle_local_attn_reshape.argtypes = [f32p, f32p] + [ctypes.c_int] * 4
        L.oracle_local_attn_reshape.restype = None
        L.oracle_block_extract_backward.argtypes = [f32p] * 5 + [ctypes.c_int] * 7
        L.oracle_block_extract_backward.restype = None
        L.oracle_local_attn_reshape_backward.argtypes = [f32p, f32p] + [ctypes.c_int] * 4
        L.oracle_local_attn_reshape_backward.restype = None
        L.oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def face_inv(faces: np.ndarray, image_size: int) -> np.ndarray:
    faces = np.ascontiguousarray(faces, np.float32).reshape(-1, 9)
    out = np.zeros_like(faces)
    lib().oracle_face_inv(faces, faces.shape[0], image_size, out)
    return out


def rasterize(faces: np.ndarray, image_size: int = 256, near: float = 0.1, far: float = 100.0,
              flip_y: bool = True, return_depth: bool = True):
    """faces (B,F,3,3) f32 -> fim (B,is,is) i32, wim (B,is,is,3) f32, depth (B,is,is) f32."""
    faces = np.ascontiguousarray(faces, np.float32)
    B, F = faces.shape[:2]
    fim = np.empty((B, image_size, image_size), np.int32)
    wim = np.empty((B, image_size, image_size, 3), np.float32)
    depth = np.empty((B, image_size, image_size), np.float32) if return_depth else None
    lib().oracle_rasterize(faces, B, F, image_size, near, far, int(flip_y), fim, wim,
                           depth.ctypes.data if depth is not None else None)
    return fim, wim, depth


def block_extract(src: np.ndarray, flow: np.ndarray, k: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.float32)
    flow = np.ascontiguousarray(flow, np.float32)
    B, C, Hs, Ws = src.shape
    _, two, Hf, Wf = flow.shape
    assert two == 2
    out = np.empty((B, C, k * Hf, k * Wf), np.float32)
    lib().oracle_block_extract(src, flow, out, B, C, Hs, Ws, Hf, Wf, k)
    return out


def block_extract_backward(src: np.ndarray, flow: np.ndarray, grad_out: np.ndarray, k: int):
    """block_extractor_kernel.cu:86-166 -> (grad_src, grad_flow), starting from zeros like block_extractor.py:36-37."""
    src = np.ascontiguousarray(src, np.float32)
    flow = np.ascontiguousarray(flow, np.float32)
    grad_out = np.ascontiguousarray(grad_out, np.float32)
    B, C, Hs, Ws = src.shape
    _, _, Hf, Wf = flow.shape
    gs, gf = np.zeros_like(src), np.zeros_like(flow)
    lib().oracle_block_extract_backward(src, flow, grad_out, gs, gf, B, C, Hs, Ws, Hf, Wf, k)
    return gs, gf


def local_attn_reshape_backward(grad_out: np.ndarray, k: int) -> np.ndarray:
    grad_out = np.ascontiguousarray(grad_out, np.float32)
    B, _, Ho, Wo = grad_out.shape
    gin = np.zeros((B, k * k, Ho // k, Wo // k), np.float32)
    lib().oracle_local_attn_reshape_backward(grad_out, gin, B, k, Ho // k, Wo // k)
    return gin


def local_attn_reshape(x: np.ndarray, k: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    B, kk, H, W = x.shape
    assert kk == k * k
    out = np.empty((B, 1, k * H, k * W), np.float32)
    lib().oracle_local_attn_reshape(x, out, B, k, H, W)
    return out
