"""Config 4: condition rasterizer alone, N meshes (F = 13776) -> 256x256 fim/wim.  CUDA-event timing and HBM roofline.
(bench.py's `rasterizer` extra also checks it bit-for-bit against, and times, the reference's own kernel.)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hoig_b200 import ops, synth  # noqa: E402
from hoig_b200.renderer import EYE_Z  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
if len(sys.argv) > 2:
    from hoig_b200 import _lib
    _lib.load().hoig_set_rasterizer_band_pixels(int(sys.argv[2]))
CH = 256  # distinct pose pairs generated on host, tiled to N on device
sc = synth.make_scene(CH // 2, seed=0, obj_faces=12238)
nv = sc.n_verts
verts = torch.cat([sc.verts_src[:, :nv], sc.verts_ref[:, :nv]], 0).contiguous().cuda()
cam = torch.cat([sc.cam, sc.cam], 0).cuda()
fidx = sc.faces_idx.cuda()
faces = ops.project_faces(verts, cam, fidx, EYE_Z)
faces = faces.repeat(N // CH, 1, 1, 1).contiguous()
F = faces.shape[1]
print("faces", tuple(faces.shape), f"{faces.numel() * 4 / 1e9:.2f} GB")
for _ in range(2):
    fim, wim, depth = ops.rasterize(faces, 256, return_depth=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    fim, wim, depth = ops.rasterize(faces, 256, return_depth=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bytes_per_mesh = F * 36 + 65536 * (4 + 12 + 4)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
gbs = N * bytes_per_mesh / ms / 1e6
print(json.dumps({"meshes": N, "faces": F, "ms": ms, "meshes_per_s": N / ms * 1e3, "algorithmic_bytes_per_mesh": bytes_per_mesh,
                  "achieved_GBs": gbs, "hbm_peak_GBs": peak, "frac": gbs / peak, "covered_frac": (fim >= 0).float().mean().item()}))
