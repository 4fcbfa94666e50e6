"""Generates tests/golden/geometry_stage_r.npz by running the UNMODIFIED reference stage-R code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_geometry.py

What runs -- the reference's own source, imported from /root/reference/HOIG_HOv3, nothing restated:
  * ``HandRecoveryFlow.forward``                          models/trainer.py:46-145 (the whole per-sample loop + channel algebra)
  * ``MANORenderer.render_fim_wim / encode_fim / encode_sem / cal_bc_transform / get_texture_backward_warp /
    sample_from_texture_dense``                          utils/nmr.py:496-513, 567-595, 874-968, 973-1058, 1068-1100
  * ``orthographic_proj_withz_idrot``                     utils/nmr.py:109-140
  * ``nr.look_at``, ``nr.vertices_to_faces``              thirdparty/neural_renderer/neural_renderer/{look_at,vertices_to_faces}.py
  * ``util.morph``                                        utils/util.py:142-158
How it is made to run here:
  * import stubs for h5py / smplx / tensorboardX and the compiled extension modules;
  * ``MANORenderer`` and ``HandRecoveryFlow`` are created with ``__new__`` (their ``__init__`` need the MANO assets) and given
    synthetic buffers of the reference's names and shapes (``faces_<obj>``, ``map_fn_<obj>``, ``sem_full_<obj>``, ``fim_uv_<obj>``,
    ``wim_uv_<obj>``, ``faces_uv_coord_<obj>``, ``obj_tex_img_<obj>``) built by ``hoig_b200.synth``;
  * ``Tensor.cuda`` is neutralised (nmr.py:124, 932-939, 1013-1020, 1053 call it unconditionally);
  * the one CUDA-only call, ``nr.rasterize_face_index_map_and_weight_map`` (rasterize.py:543-571 allocates torch.cuda tensors),
    is served by the C oracle, which is bit-equal to the reference's own rasterizer kernels on the GPU
    (tests/test_gpu_rasterizer.py) -- same init values, same flip.
The wrapped renderer methods record their return values, so the fixture pins every stage-R function, not only the final
generator inputs.

Fixture layout: piecewise-constant tensors (masks, one-hot maps, encoded UV maps, fim) are stored in full; noise-like tensors
(anything multiplied by the random source image, textures) as a strided sample (``*_s``, stride 4) plus float64 sums (``*_sum``);
the (B,F,3,3) projected faces and T maps in full.  Inputs are regenerated from ``hoig_b200.synth`` seeds (checksums stored).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/HOIG_HOv3"

from tests.geometry_inputs import B, OBJ_SLOT, STRIDE, checksum, scene_and_tables  # noqa: E402


def import_reference():
    import oracle

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return object

    for m in ["h5py", "smplx", "smplx.lbs", "smplx.utils", "smplx.vertex_ids", "smplx.vertex_joint_selector", "tensorboardX",
              "neural_renderer.cuda", "neural_renderer.cuda.rasterize", "neural_renderer.cuda.load_textures",
              "neural_renderer.cuda.create_texture_image", "block_extractor_cuda", "local_attn_reshape_cuda"]:
        sys.modules.setdefault(m, _Any(m))
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "thirdparty", "neural_renderer"))
    import neural_renderer as nr
    import models.trainer as trainer
    import utils.nmr as nmr

    def rasterize_fim_wim(faces, image_size=256, anti_aliasing=False, near=0.1, far=100.0, eps=1e-4):
        assert not anti_aliasing
        fim, wim, _ = oracle.rasterize(faces.detach().numpy(), image_size, near, far, flip_y=True, return_depth=False)
        return torch.from_numpy(fim), torch.from_numpy(wim)

    nr.rasterize_face_index_map_and_weight_map = rasterize_fim_wim       # nmr.py calls it through the module attribute
    torch.Tensor.cuda = lambda self, *a, **k: self
    return trainer, nmr


class _Recorder:
    """Wraps bound methods of the renderer; keeps every call's return value."""

    def __init__(self, obj, names):
        self.calls = {n: [] for n in names}
        for n in names:
            fn = getattr(obj, n)

            def wrapped(*a, _fn=fn, _n=n, **k):
                out = _fn(*a, **k)
                self.calls[_n].append(out)
                return out

            setattr(obj, n, wrapped)


def run_reference(inp):
    trainer, nmr = import_reference()
    sc = inp["scene"]
    name = trainer.OBJNAMES[OBJ_SLOT]
    r = nmr.MANORenderer.__new__(nmr.MANORenderer)
    torch.nn.Module.__init__(r)
    r.image_size = 256
    r.proj_func = nmr.orthographic_proj_withz_idrot
    r.eye = [0, 0, -(1. / np.tan(np.radians(30)) + 1)]                     # nmr.py:357 with the default viewing_angle=30
    r.register_buffer("faces_" + name, sc.faces_idx.clone().int())
    r.register_buffer("map_fn_" + name, inp["map_fn"])
    r.register_buffer("sem_full_" + name, inp["sem_full"])
    r.register_buffer("fim_uv_" + name, inp["fim_uv"][None])
    r.register_buffer("wim_uv_" + name, inp["wim_uv"][None])
    r.register_buffer("faces_uv_coord_" + name, inp["coord"][None])
    r.register_buffer("obj_tex_img_" + name, inp["obj_tex"])
    rec = _Recorder(r, ["render_fim_wim", "encode_fim", "encode_sem", "cal_bc_transform", "get_texture_backward_warp",
                        "sample_from_texture_dense"])

    class _HMR:
        def __init__(self):
            self.n = 0

        def get_details(self, mano):
            self.n += 1
            verts = sc.verts_src if self.n == 1 else sc.verts_ref
            return {"objName": torch.full((B,), OBJ_SLOT, dtype=torch.long), "cam": sc.cam.clone(), "verts": verts.clone()}

    flow = trainer.HandRecoveryFlow.__new__(trainer.HandRecoveryFlow)
    torch.nn.Module.__init__(flow)
    flow._opt = types.SimpleNamespace(bg_both=False)
    flow._hmr = _HMR()
    flow._render = r
    with torch.no_grad():
        outs = flow.forward(inp["src_img"].clone(), inp["ref_img"].clone(), None, None)
    return outs, rec.calls


def put(rec, key, t, full):
    t = t.detach()
    if full:
        rec[key] = t.numpy()
    else:
        rec[key + "_s"] = t[..., ::STRIDE, ::STRIDE].contiguous().numpy() if t.dim() == 4 and t.shape[1] <= 16 else t.numpy()
        rec[key + "_sum"] = np.array([t.double().sum().item(), t.double().abs().sum().item()])


def main():
    inp = scene_and_tables()
    sc = inp["scene"]
    outs, calls = run_reference(inp)
    names = ["input_G_src_bg", "input_G_tsf_bg", "input_G_src_obj", "input_G_tsf_obj", "input_G_src_hand", "input_G_ref_hand",
             "T_hand", "src_crop_mask_bg", "ref_crop_mask_bg", "src_crop_mask_hand", "ref_crop_mask_hand"]
    rec = {"input_checksum": np.array(checksum(sc.faces_idx, sc.verts_src, sc.verts_ref, sc.cam, inp["map_fn"], inp["sem_full"],
                                               inp["fim_uv"], inp["wim_uv"], inp["coord"], inp["src_img"], inp["obj_tex"])),
           "n_faces": np.array(sc.n_faces)}
    o = dict(zip(names, outs))
    assert o["input_G_tsf_bg"] is None
    # final outputs of HandRecoveryFlow.forward (trainer.py:144-145)
    put(rec, "out_input_G_src_bg_rgb", o["input_G_src_bg"][:, :3], False)
    put(rec, "out_input_G_src_bg_mask", o["input_G_src_bg"][:, 3:].to(torch.uint8), True)
    for k in ("src_obj", "tsf_obj"):
        put(rec, f"out_input_G_{k}_rgb", o[f"input_G_{k}"][:, :3], False)
        put(rec, f"out_input_G_{k}_cond", o[f"input_G_{k}"][:, 3:6], True)
        put(rec, f"out_input_G_{k}_seg", o[f"input_G_{k}"][:, 6:].to(torch.uint8), True)
    for k in ("src_hand", "ref_hand"):
        put(rec, f"out_input_G_{k}_rgb", o[f"input_G_{k}"][:, :3], False)
        put(rec, f"out_input_G_{k}_cond", o[f"input_G_{k}"][:, 3:], True)
    put(rec, "out_T_hand", o["T_hand"], True)
    for k in names[7:]:
        put(rec, "out_" + k, o[k].to(torch.uint8), True)
    # per-function records, one entry per sample (the reference loops over the batch)
    for tag, idx in (("src", 0), ("ref", 1)):
        faces = torch.cat([calls["render_fim_wim"][2 * i + idx][0] for i in range(B)])
        fim = torch.cat([calls["render_fim_wim"][2 * i + idx][1] for i in range(B)])
        wim = torch.cat([calls["render_fim_wim"][2 * i + idx][2] for i in range(B)])
        # src faces come back with y negated in place by trainer.py:67-68 (xy view of the same storage): undo for the record
        if tag == "src":
            faces = faces.clone()
            faces[..., 1] *= -1
        put(rec, f"faces_{tag}", faces, True)
        put(rec, f"fim_{tag}", fim, True)
        cov = fim != -1
        rec[f"wim_{tag}_covered"] = wim[cov].numpy()                          # (n_covered, 3): wim is 0 elsewhere
        cond = torch.cat([calls["encode_fim"][2 * i + idx][0] for i in range(B)])
        sem = torch.cat([calls["encode_sem"][2 * i + idx][0] for i in range(B)])
        put(rec, f"cond_{tag}", cond, True)
        put(rec, f"sem_{tag}", sem.to(torch.uint8), True)
        Tt = torch.cat([calls["sample_from_texture_dense"][2 * i + (1 - idx)] for i in range(B)])   # called ref first (trainer.py:84,86)
        put(rec, f"T_tex_{tag}", Tt, True)
    T = torch.cat([calls["cal_bc_transform"][i][0] for i in range(B)])
    O = torch.cat([calls["cal_bc_transform"][i][1] for i in range(B)])
    put(rec, "bc_T", T, True)
    put(rec, "bc_O", O.to(torch.uint8), True)
    tex = torch.cat([calls["get_texture_backward_warp"][i] for i in range(B)])
    put(rec, "texture", tex, False)
    rec["texture_s"] = tex[:, :, ::STRIDE, ::STRIDE].contiguous().numpy()
    path = os.path.join(HERE, "geometry_stage_r.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", {k: v.shape for k, v in rec.items() if hasattr(v, "shape") and v.ndim > 1})
    for tag in ("src", "ref"):
        print(tag, "covered fraction", float((rec[f"fim_{tag}"] != -1).mean()))


if __name__ == "__main__":
    main()
