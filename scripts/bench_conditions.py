"""Stage R end to end at batch B (default 64): projection, two rasterizations, condition maps, masks, the dense correspondence T
and the UV-texture warp (R0-R8), all batched through the C ABI -- the work `HandRecoveryFlow.forward` (models/trainer.py:63-145)
does with a per-sample Python loop.  Prints one JSON line with per-stage times and the HBM roofline fraction of the whole stage.

    python scripts/bench_conditions.py [B]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hoig_b200 import ops, renderer, synth  # noqa: E402


def uv_atlas(scene, dev):
    coord, fim_uv, wim_uv = synth.uv_atlas(scene, lambda tri: ops.rasterize(tri.to(dev), 256, flip_y=False))
    return coord.to(dev), fim_uv.to(dev), wim_uv.to(dev)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = "cuda"
    sc = synth.make_scene(B, seed=0)
    faces_idx = sc.faces_idx.to(dev)
    cam, vs, vr = sc.cam.to(dev), sc.verts_src.to(dev), sc.verts_ref.to(dev)
    map_fn, sem = sc.map_fn.to(dev), sc.sem_full.to(dev)
    g = torch.Generator().manual_seed(0)
    src_img = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    obj_tex = torch.rand(256, 256, 3, generator=g).to(dev)
    coord, fim_uv, wim_uv = uv_atlas(sc, dev)
    F = faces_idx.shape[0]

    stages = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        stages.setdefault(name, []).append((e0, e1))
        return out

    def step():
        fs, fim_s, wim_s = timed("R0-R3 project + rasterize (src)", lambda: renderer.render_fim_wim_batched(cam, vs, faces_idx))
        fr, fim_r, wim_r = timed("R0-R3 project + rasterize (ref)", lambda: renderer.render_fim_wim_batched(cam, vr, faces_idx))
        tex = timed("R8 texture atlas (backward warp + grid_sample + compose)",
                    lambda: renderer.texture_backward_warp(src_img, fs, fim_s, fim_uv, wim_uv, obj_tex))
        r_ref = timed("R8 re-render (ref pose)", lambda: renderer.render_from_texture(tex, fim_r, wim_r, coord))
        r_src = timed("R8 re-render (src pose)", lambda: renderer.render_from_texture(tex, fim_s, wim_s, coord))
        return timed("R4-R7 condition maps, masks, T, input assembly (one launch)",
                     lambda: renderer.condition_inputs_fused(src_img, fs, fim_s, fim_r, wim_r, map_fn, sem, r_src, r_ref))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    stages.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    t0.record()
    for _ in range(iters):
        out = step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / iters
    per = {k: sum(a.elapsed_time(b) for a, b in v) / iters for k, v in stages.items()}
    # algorithmic bytes per sample: two rasterizations + atlas (T, O, texture) + two renders + condition tensors
    rast = 2 * (F * 36 + 65536 * 16)
    atlas = 256 * 640 * (4 + 12) / B + 256 * 640 * (8 + 4 + 12 + 12) + 3 * 65536 * 4
    renders = 2 * (65536 * (4 + 12 + 8) + 2 * 3 * 65536 * 4)
    conds = 2 * 65536 * 4 * (3 + 15 + 1 + 1 + 1) + 65536 * (16 + 8)
    bytes_per_sample = rast + atlas + renders + conds
    peak = 6553.0
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    inputs, _masks = out
    cover = (inputs["T"][..., 0] > -1.5).float().mean().item()
    print(json.dumps({"stage": "R0-R8 condition stage (batched HandRecoveryFlow.forward)", "batch": B, "faces": F, "ms_per_batch": ms,
                      "samples_per_s": B / ms * 1e3, "per_stage_ms": {k: round(v, 3) for k, v in per.items()},
                      "algorithmic_bytes_per_sample": int(bytes_per_sample), "achieved_GBs": bytes_per_sample * B / ms / 1e6,
                      "hbm_peak_GBs": peak, "frac": bytes_per_sample * B / ms / 1e6 / peak, "T_valid_fraction": cover}))


if __name__ == "__main__":
    main()
