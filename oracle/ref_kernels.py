"""Loader for the reference's OWN CUDA kernels built by oracle/build_ref.py into oracle/_ref -- TEST INFRASTRUCTURE ONLY
(the same-toolchain, same-GPU oracle of SURVEY.md section 8c; used by tests/ and by bench.py's checker legs)."""
import importlib.util
import os

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def load_ref(name):
    """Imports oracle/_ref/<name>.so (a torch extension) or returns None when it was not built."""
    path = os.path.join(REF_DIR, name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    # the reference's rasterizer hard-codes its pybind module name (rasterize_cuda.cpp:194)
    init = "rasterize" if name == "ref_rasterize_cuda" else name
    spec = importlib.util.spec_from_file_location(init, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_rasterize(mod, faces, is_, near=0.1, far=100.0):
    """rasterize.py:50-52 allocation + rasterize_cuda.cpp:70 call + rasterize.py:335-338 flip."""
    import torch
    B, F = faces.shape[:2]
    dev = faces.device
    fim = torch.full((B, is_, is_), -1, dtype=torch.int32, device=dev)
    wim = torch.zeros(B, is_, is_, 3, device=dev)
    depth = torch.full((B, is_, is_), far, device=dev)
    finv_map = torch.zeros(1, device=dev)
    finv = torch.zeros(B, F, 3, 3, device=dev)
    mod.forward_face_index_map(faces.clone(), fim, wim, depth, finv_map, finv, is_, near, far, 0, 0, 0)
    torch.cuda.synchronize()
    return torch.flip(fim, dims=(1,)), torch.flip(wim, dims=(1,)), torch.flip(depth, dims=(1,)), finv
