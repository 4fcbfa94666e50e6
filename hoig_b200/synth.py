"""Deterministic synthetic inputs for tests and bench (no dataset, no MANO assets).

The reference needs MANO assets and HO3D annotations that are not shipped
(SURVEY.md section 8c/8d); this module builds stand-ins of exactly the same
shapes and value ranges:

* a "hand" surface with the MANO right-hand counts, 778 vertices / 1538 faces
  (any triangulated disk with 16 boundary and 762 interior vertices has that
  face count: F = 2*V_int + V_b - 2),
* a closed "object" surface whose face indices are offset by 778
  (utils/nmr.py:286) and whose vertices are zero-padded to 7866 rows
  (data/hov3_dataset.py:246-248),
* an HO3D-like camera row ``cam`` (B,15) = K (3x3) | affine (2x3)
  (utils/nmr.py:109-140),
* per-face tables ``map_fn`` (F+1,3) / ``sem_full`` (F+1,1) with the
  background row last (utils/nmr.py:297-330, utils/mesh.py:393-396),
* random generator inputs of the shapes ``HandRecoveryFlow.forward`` emits
  (models/trainer.py:127-136).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

N_HAND_V, N_HAND_F = 778, 1538
N_OBJ_V_PAD = 7866


def _fib_disk(n: int, rmax: float) -> np.ndarray:
    i = np.arange(n, dtype=np.float64) + 0.5
    r = rmax * np.sqrt(i / n)
    t = i * math.pi * (3.0 - math.sqrt(5.0))
    return np.stack([r * np.cos(t), r * np.sin(t)], 1)


def _fib_sphere(n: int) -> np.ndarray:
    i = np.arange(n, dtype=np.float64) + 0.5
    z = 1.0 - 2.0 * i / n
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    t = i * math.pi * (3.0 - math.sqrt(5.0))
    return np.stack([r * np.cos(t), r * np.sin(t), z], 1)


def hand_mesh():
    """778 v / 1538 f open 'thimble' (disk topology, 16 boundary edges), metres."""
    from scipy.spatial import Delaunay

    tb = np.arange(16) * (2 * math.pi / 16)
    boundary = np.stack([np.cos(tb), np.sin(tb)], 1)
    interior = _fib_disk(N_HAND_V - 16, 0.97 * math.cos(math.pi / 16))
    p2 = np.concatenate([boundary, interior], 0)
    tri = Delaunay(p2).simplices.astype(np.int64)
    a, b, c = p2[tri[:, 0]], p2[tri[:, 1]], p2[tri[:, 2]]
    ccw = ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])) > 0
    tri[~ccw] = tri[~ccw][:, [0, 2, 1]]
    assert tri.shape[0] == N_HAND_F, tri.shape
    r2 = (p2 ** 2).sum(1)
    # dome: wrist ring at z=0, finger tip at z=0.16; slightly flattened in y
    v = np.stack([0.045 * p2[:, 0], 0.03 * p2[:, 1], 0.16 * np.sqrt(np.maximum(0.0, 1.0 - r2))], 1)
    order = np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))
    return v.astype(np.float32), tri[order].astype(np.int32)


def object_mesh(n_faces: int = 12238):
    """Closed ellipsoid with exactly ``n_faces`` (even) faces; V = n_faces/2 + 2."""
    from scipy.spatial import ConvexHull

    assert n_faces % 2 == 0
    nv = n_faces // 2 + 2
    assert nv <= N_OBJ_V_PAD
    p = _fib_sphere(nv)
    tri = ConvexHull(p).simplices.astype(np.int64)
    a, b, c = p[tri[:, 0]], p[tri[:, 1]], p[tri[:, 2]]
    out = (np.cross(b - a, c - a) * (a + b + c)).sum(1) > 0
    tri[~out] = tri[~out][:, [0, 2, 1]]
    assert tri.shape[0] == n_faces, tri.shape
    v = p * np.array([0.04, 0.075, 0.03])
    order = np.lexsort((tri[:, 2], tri[:, 1], tri[:, 0]))
    return v.astype(np.float32), tri[order].astype(np.int32)


def _rot(rng: np.random.Generator, max_angle: float) -> np.ndarray:
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = rng.uniform(-max_angle, max_angle)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


@dataclass
class Scene:
    faces_idx: torch.Tensor   # (F,3) int32, object faces offset by 778
    verts_src: torch.Tensor   # (B, 778+7866, 3) f32, OpenGL coords (object rows zero padded)
    verts_ref: torch.Tensor
    cam: torch.Tensor         # (B,15)
    n_verts: int              # valid rows = faces_idx.max()+1
    map_fn: torch.Tensor      # (F+1,3)
    sem_full: torch.Tensor    # (F+1,1)

    @property
    def n_faces(self) -> int:
        return int(self.faces_idx.shape[0])


def default_cam(B: int) -> torch.Tensor:
    K = np.array([[614.0, 0, 320.0], [0, 614.0, 240.0], [0, 0, 1.0]], np.float32)
    s = 256.0 / 480.0
    A = np.array([[s, 0, -80.0 * s], [0, s, 0.0]], np.float32)
    row = np.concatenate([K.reshape(-1), A.reshape(-1)])
    return torch.from_numpy(np.tile(row[None], (B, 1)).astype(np.float32))


def face_tables(faces_idx: np.ndarray, n_obj_id: int = 7):
    """Synthetic ``map_fn`` (UV-barycentre + segment flag) and ``sem_full`` tables.

    Layout follows utils/nmr.py:297-330: hand rows first (u,v in [0,1], third
    column 0), object rows with u offset by +1.5 (object slot 0), last row the
    background ``[0,0,1]`` resp. semantic id 0; hand part ids 1..6 by face-index
    range, object id ``n_obj_id`` (7..15).
    """
    F = faces_idx.shape[0]
    rng = np.random.default_rng(1234)
    map_fn = np.zeros((F + 1, 3), np.float32)
    uv = rng.random((F, 2), dtype=np.float32)
    map_fn[:F, :2] = uv
    map_fn[N_HAND_F:F, 0] += 1.5
    map_fn[F] = (0.0, 0.0, 1.0)
    sem = np.zeros((F + 1, 1), np.float32)
    sem[:N_HAND_F, 0] = 1 + (np.arange(N_HAND_F) * 6 // N_HAND_F)
    sem[N_HAND_F:F, 0] = n_obj_id
    return torch.from_numpy(map_fn), torch.from_numpy(sem)


def make_scene(B: int, seed: int = 0, obj_faces: int = 12238) -> Scene:
    """B seeded (src, ref) pose pairs of the hand+object assembly."""
    hv, hf = hand_mesh()
    ov, of = object_mesh(obj_faces)
    faces_idx = np.concatenate([hf, of + N_HAND_V], 0).astype(np.int32)
    base = np.concatenate([hv - np.array([0, 0, 0.08], np.float32),
                           ov + np.array([0.03, 0.0, 0.02], np.float32)], 0).astype(np.float64)
    nv = base.shape[0]
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(2):
        vs = np.zeros((B, N_HAND_V + N_OBJ_V_PAD, 3), np.float32)
        for b in range(B):
            R = _rot(rng, math.pi)
            t = np.array([rng.uniform(-0.06, 0.06), rng.uniform(-0.06, 0.06), rng.uniform(-0.6, -0.45)])
            vs[b, :nv] = (base @ R.T + t).astype(np.float32)
        out.append(torch.from_numpy(vs))
    map_fn, sem = face_tables(faces_idx)
    return Scene(torch.from_numpy(faces_idx), out[0], out[1], default_cam(B), nv, map_fn, sem)


def pose_batch(scene: Scene, n: int, seed: int = 0, device="cpu", z_range=(-0.6, -0.45), xy_range: float = 0.06):
    """``n`` DISTINCT seeded rigid poses of the scene's hand+object assembly, generated with torch on ``device`` (so that
    thousands of meshes need no Python loop): returns verts (n, n_verts, 3) f32 in OpenGL coordinates.  ``z_range`` sets the
    camera distance, i.e. the fraction of the frame the meshes cover (HO3D-like default ~7 %, (-0.3, -0.25) ~25 %)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    base = scene.verts_src[0, : scene.n_verts].double()
    # undo the first pose's translation roughly: centre the assembly, then apply fresh rotations / translations
    base = base - base.mean(0, keepdim=True)
    axis = torch.randn(n, 3, generator=g, dtype=torch.float64)
    axis = axis / axis.norm(dim=1, keepdim=True)
    ang = (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
    K = torch.zeros(n, 3, 3, dtype=torch.float64)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
    R = torch.eye(3, dtype=torch.float64)[None] + torch.sin(ang)[:, None, None] * K + (1 - torch.cos(ang))[:, None, None] * (K @ K)
    t = torch.stack([(torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1) * xy_range,
                     (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1) * xy_range,
                     z_range[0] + torch.rand(n, generator=g, dtype=torch.float64) * (z_range[1] - z_range[0])], 1)
    R, t, base = R.to(device), t.to(device), base.to(device)
    return (torch.einsum("vk,nik->nvi", base, R) + t[:, None, :]).float().contiguous()


def uv_atlas(scene: Scene, rasterize, n_hand_faces: int = N_HAND_F):
    """Synthetic UV atlas in the reference's layout (utils/nmr.py:359-401: hand atlas | 128-px gap | object atlas, 256 x 640).

    Per-vertex UVs come from a planar (hand) / spherical (object) projection of the rest pose; ``rasterize(tri)`` maps
    ``tri`` (1,Fp,3,3) f32 (UV triangles at z = 1) to ``(fim (1,256,256) int32, wim (1,256,256,3) f32)`` WITHOUT the vertical
    flip -- the caller supplies the rasterizer (the CUDA op in bench / product code, the C oracle in CPU tests).
    Returns ``(faces_uv_coord (F,3,2), fim_uv (256,640) int32, wim_uv (256,640,3))``: the tables the reference registers as
    ``faces_uv_coord_<obj>[0]``, ``fim_uv_<obj>[0]``, ``wim_uv_<obj>[0]``; atlas coordinates are in grid_sample's
    align_corners=True convention like the reference's ``(uv - mean) * scale`` (nmr.py:393-395)."""
    faces_idx = scene.faces_idx.long()
    v = scene.verts_src[0, : scene.n_verts].double()
    nh = N_HAND_V
    uv = torch.zeros(v.shape[0], 2, dtype=torch.float64)
    h = v[:nh, :2] - v[:nh, :2].mean(0)
    uv[:nh] = 0.9 * h / h.abs().max()
    o = v[nh:] - v[nh:].mean(0)
    o = o / o.norm(dim=1, keepdim=True).clamp_min(1e-9)
    uv[nh:, 0] = 0.9 * torch.atan2(o[:, 1], o[:, 0]) / math.pi
    uv[nh:, 1] = 0.9 * (2 * torch.acos(o[:, 2].clamp(-1, 1)) / math.pi - 1)
    fuv = uv[faces_idx].float()                                   # (F,3,2) local [-1,1]^2 coordinates of each part
    fim_uv = torch.full((256, 640), -1, dtype=torch.int32)
    wim_uv = torch.zeros(256, 640, 3)
    for lo, hi, x0 in ((0, n_hand_faces, 0), (n_hand_faces, fuv.shape[0], 384)):
        tri = torch.cat([fuv[lo:hi], torch.ones(hi - lo, 3, 1)], 2)[None].contiguous()             # z = 1: inside the frustum
        fim, wim = rasterize(tri)
        fim, wim = fim.cpu(), wim.cpu()
        fim_uv[:, x0:x0 + 256] = torch.where(fim[0] >= 0, fim[0] + lo, fim[0])
        wim_uv[:, x0:x0 + 256] = wim[0]
    px = (fuv[..., 0] + 1) / 2 * 255
    px[n_hand_faces:] += 384
    py = (fuv[..., 1] + 1) / 2 * 255
    coord = torch.stack([px / 639 * 2 - 1, py / 255 * 2 - 1], -1).contiguous()
    return coord, fim_uv.contiguous(), wim_uv.contiguous()


def generator_inputs(B: int, seed: int = 0, size: int = 256, img_cond_dim: int = 3,
                     obj_cond_dim: int = 12, device="cpu", with_holes: bool = True):
    """Random tensors with the shapes/ranges ``Generator.forward`` receives
    (models/trainer.py:377-393): returns the kwargs dict of that call."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*s):
        return torch.rand(*s, generator=g) * 2 - 1

    T = rnd(B, size, size, 2)
    if with_holes:  # invalid correspondences carry the -2 sentinel (utils/nmr.py:884)
        hole = torch.rand(B, size, size, 1, generator=g) < 0.3
        T = torch.where(hole, torch.full_like(T, -2.0), T)
    d = dict(
        bg_inputs=rnd(B, 4, size, size),
        src_obj_inputs=rnd(B, 3, size, size),
        tsf_obj_inputs=rnd(B, 3, size, size),
        src_hand_inputs=rnd(B, 3, size, size),
        tsf_hand_inputs=rnd(B, 3, size, size),
        T=T,
        src_obj_conds=torch.rand(B, obj_cond_dim, size, size, generator=g),
        src_hand_conds=torch.rand(B, img_cond_dim, size, size, generator=g),
        tsf_obj_conds=torch.rand(B, obj_cond_dim, size, size, generator=g),
        tsf_hand_conds=torch.rand(B, img_cond_dim, size, size, generator=g),
        src_armask=(torch.rand(B, 1, size, size, generator=g) > 0.5).float(),
        tsf_armask=(torch.rand(B, 1, size, size, generator=g) > 0.5).float(),
    )
    return {k: v.to(device) for k, v in d.items()}
