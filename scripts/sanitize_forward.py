"""One GeneratorB200 forward + composite + stage R for compute-sanitizer runs (memcheck / racecheck / synccheck).
generator_spade_attn, conv_dim 64, repeat_num 2, batch 2, 256x256 (so that the row-halo / vertical-halo / CTA-pair / dual-pipeline /
resident-weight / tap-skip modes of conv_umma and conv_halo, attn_combine and the norm / layout kernels all run), default fp16."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hoig_b200 import ops, renderer, synth  # noqa: E402
from hoig_b200.generator import composite, create  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = dict(bg_dim=8, img_dim=3, obj_dim=3, img_cond_dim=3, obj_cond_dim=12, conv_dim=64, repeat_num=2)
g = create("generator_spade_attn", **cfg).cuda().eval()
g.auto_graph = False
inp = {k: v.cuda() for k, v in synth.generator_inputs(2, seed=1, size=size).items()}
o = g(**inp)
img = composite(o[1], o[6], o[7], o[8], o[9])
sc = synth.make_scene(2, seed=0, obj_faces=2000)
faces, fim, wim = renderer.render_fim_wim_batched(sc.cam.cuda(), sc.verts_src[:, :sc.n_verts].cuda().contiguous(), sc.faces_idx.cuda())
torch.cuda.synchronize()
print("forward ok", float(img.abs().mean()), int((fim >= 0).sum()))
