"""Deterministic stage-R inputs shared by tests/golden/make_golden_geometry.py (which runs the unmodified reference on them) and
the tests that compare the oracle restatement and the CUDA path with the resulting fixture (tests/golden/geometry_stage_r.npz)."""
import hashlib

import numpy as np
import torch

B, SEED, OBJ_FACES, OBJ_SLOT = 2, 3, 2000, 4        # OBJ_SLOT: index into trainer.OBJNAMES ('011_banana')
STRIDE = 4


def scene_and_tables():
    """Deterministic inputs shared by this script and the tests (CPU tensors)."""
    import oracle
    from hoig_b200 import synth

    sc = synth.make_scene(B, seed=SEED, obj_faces=OBJ_FACES)

    def rast(tri):
        fim, wim, _ = oracle.rasterize(tri.numpy(), 256, flip_y=False, return_depth=False)
        return torch.from_numpy(fim), torch.from_numpy(wim)

    coord, fim_uv, wim_uv = synth.uv_atlas(sc, rast)
    g = torch.Generator().manual_seed(SEED)
    src_img = torch.rand(B, 3, 256, 256, generator=g) * 2 - 1
    ref_img = torch.rand(B, 3, 256, 256, generator=g) * 2 - 1
    obj_tex = torch.rand(256, 256, 3, generator=g) * 2 - 1
    # the reference's map_fn offsets object slot i by 1.5 * (i + 1) in u (nmr.py:325); synth builds slot 0
    map_fn = sc.map_fn.clone()
    map_fn[synth.N_HAND_F:-1, 0] += 1.5 * OBJ_SLOT
    sem = sc.sem_full.clone()
    sem[synth.N_HAND_F:-1, 0] = OBJ_SLOT + 7                     # nmr.py:311
    return dict(scene=sc, coord=coord, fim_uv=fim_uv, wim_uv=wim_uv, src_img=src_img, ref_img=ref_img, obj_tex=obj_tex,
                map_fn=map_fn, sem_full=sem)


def checksum(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.numpy()).tobytes())
    return h.hexdigest()


