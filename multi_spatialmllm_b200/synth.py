"""Seeded synthetic stand-ins for ScanNet scenes and TAPVid-3D videos.

ScanNet / TAPVid-3D data are not redistributable and are absent from both the
build container and the GPU box, so every BASELINE.json config runs on scenes
made here.  Shapes and conventions follow the reference's
on-disk data: axis-aligned vertices that are float32-representable stored as
float64 (batch_load_scannet_data.py:66-69), camera->world poses rounded to six
decimals like the ``%f`` pose text files (extract_posed_images.py:139-142),
4x4 intrinsics, 1296x968 colour / 640x480 uint16-mm depth.

Host-side NumPy only; no geometry of the hot path is evaluated here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

IMAGE_H, IMAGE_W = 968, 1296
DEPTH_H, DEPTH_W = 480, 640
SCANNET_K = np.array([[1170.187988, 0.0, 647.75, 0.0],
                      [0.0, 1170.187988, 483.75, 0.0],
                      [0.0, 0.0, 1.0, 0.0],
                      [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)


@dataclass
class SynthScene:
    scene_id: str
    points: np.ndarray            # (N,3) float32, axis-aligned frame
    instance_ids: np.ndarray      # (N,) uint16, 1-indexed object id, 0 = none
    num_objects: int
    axis_align: np.ndarray        # (4,4) float64
    c2w: np.ndarray               # (F,4,4) float64, raw (un-aligned) camera->world
    K: np.ndarray                 # (4,4) float64
    image_hw: tuple = (IMAGE_H, IMAGE_W)
    depth_hw: tuple = (DEPTH_H, DEPTH_W)
    depth: np.ndarray | None = None   # (F,dh,dw) uint16 once rendered
    image_ids: list = field(default_factory=list)

    @property
    def n_points(self) -> int:
        return int(self.points.shape[0])

    @property
    def n_frames(self) -> int:
        return int(self.c2w.shape[0])


def _grid_face(rng, origin, eu, ev, n, jitter):
    """~n points on a rectangle, emitted in scan-line order (mesh-like locality)."""
    lu, lv = np.linalg.norm(eu), np.linalg.norm(ev)
    nu = max(2, int(round(np.sqrt(n * lu / lv))))
    nv = max(2, int(np.ceil(n / nu)))
    a, b = np.meshgrid((np.arange(nu) + 0.5) / nu, (np.arange(nv) + 0.5) / nv, indexing="xy")
    a = a.ravel()[:n] + rng.uniform(-jitter, jitter, n) / nu
    b = b.ravel()[:n] + rng.uniform(-jitter, jitter, n) / nv
    return origin[None, :] + a[:, None] * eu[None, :] + b[:, None] * ev[None, :]


def _box_faces(lo, hi):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    return [(lo, ex, ey), (lo + ez, ex, ey), (lo, ex, ez), (lo + ey, ex, ez), (lo, ey, ez), (lo + ex, ey, ez)]


def look_at_c2w(eye, target, up=(0.0, 0.0, 1.0)):
    """Camera->world with +z forward, +x right, +y down (ScanNet/OpenCV convention)."""
    eye, target, up = np.asarray(eye, float), np.asarray(target, float), np.asarray(up, float)
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    if np.linalg.norm(r) < 1e-6:
        r = np.cross(f, np.array([0.0, 1.0, 0.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    M = np.eye(4)
    M[:3, 0], M[:3, 1], M[:3, 2], M[:3, 3] = r, d, f, eye
    return M


def look_at_c2w_batch(eye, target, up=(0.0, 0.0, 1.0)):
    """Vectorised look_at_c2w: eye, target (F,3) -> (F,4,4)."""
    eye, target = np.asarray(eye, float), np.asarray(target, float)
    up = np.broadcast_to(np.asarray(up, float), eye.shape)
    f = target - eye
    f = f / np.linalg.norm(f, axis=1, keepdims=True)
    r = np.cross(f, up)
    bad = np.linalg.norm(r, axis=1) < 1e-6
    if bad.any():
        r[bad] = np.cross(f[bad], np.array([0.0, 1.0, 0.0]))
    r = r / np.linalg.norm(r, axis=1, keepdims=True)
    d = np.cross(f, r)
    M = np.tile(np.eye(4), (eye.shape[0], 1, 1))
    M[:, :3, 0], M[:, :3, 1], M[:, :3, 2], M[:, :3, 3] = r, d, f, eye
    return M


def make_scene(seed: int, n_points: int, n_frames: int, n_objects: int = 12,
               room=(8.0, 8.0, 3.0), scene_id: str | None = None,
               shuffle_points: bool = False, orbit: bool = False,
               invalid_pose_every: int = 0) -> SynthScene:
    """Room shell (walls, floor, ceiling) + box-shaped objects, cameras inside.

    ``orbit`` yields consecutive poses on a closed path (config 1: adjacent
    pairs); otherwise poses are independent draws (configs 2/5).
    ``invalid_pose_every`` > 0 plants -inf poses like ScanNet's lost-tracking
    frames (info_handler.py:409-418) at that stride.
    """
    rng = np.random.default_rng(seed)
    rx, ry, rz = room
    n_obj_pts = int(n_points * 0.35)
    n_shell = n_points - n_obj_pts
    chunks, ids = [], []
    shell = _box_faces((0, 0, 0), room)
    areas = np.array([np.linalg.norm(np.cross(eu, ev)) for _, eu, ev in shell])
    counts = np.floor(n_shell * areas / areas.sum()).astype(int)
    counts[0] += n_shell - counts.sum()
    for (o, eu, ev), c in zip(shell, counts):
        chunks.append(_grid_face(rng, o, eu, ev, int(c), 0.35))
        ids.append(np.zeros(int(c), np.uint16))
    per_obj = np.full(n_objects, n_obj_pts // max(n_objects, 1))
    if n_objects:
        per_obj[0] += n_obj_pts - per_obj.sum()
    for k in range(n_objects):
        size = rng.uniform([0.4, 0.4, 0.4], [1.6, 1.6, 1.4])
        lo = np.array([rng.uniform(0.3, rx - 0.3 - size[0]), rng.uniform(0.3, ry - 0.3 - size[1]), 0.0])
        faces = _box_faces(lo, lo + size)
        fa = np.array([np.linalg.norm(np.cross(eu, ev)) for _, eu, ev in faces])
        fc = np.floor(per_obj[k] * fa / fa.sum()).astype(int)
        fc[0] += per_obj[k] - fc.sum()
        for (o, eu, ev), c in zip(faces, fc):
            chunks.append(_grid_face(rng, o, eu, ev, int(c), 0.35))
            ids.append(np.full(int(c), k + 1, np.uint16))
    pts = np.concatenate(chunks).astype(np.float32)
    inst = np.concatenate(ids)
    if shuffle_points:
        perm = rng.permutation(pts.shape[0])
        pts, inst = pts[perm], inst[perm]

    ang = rng.uniform(0, 2 * np.pi)
    A = np.eye(4)
    A[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    A[:3, 3] = rng.uniform(-4, 4, 3)
    A = np.round(A, 6)
    A_inv = np.linalg.inv(A)

    ctr = np.array([rx / 2, ry / 2, 1.2])
    if orbit:
        t = 2 * np.pi * np.arange(n_frames) / max(n_frames, 1)
        eye = ctr + np.stack([0.32 * rx * np.cos(t), 0.32 * ry * np.sin(t), 0.3 * np.sin(3 * t)], 1)
        tgt = ctr + np.stack([0.45 * rx * np.cos(t + 2.2), 0.45 * ry * np.sin(t + 2.2), np.full_like(t, -0.5)], 1)
    else:
        eye = rng.uniform([0.8, 0.8, 0.9], [rx - 0.8, ry - 0.8, 2.0], size=(n_frames, 3))
        tgt = rng.uniform([0.0, 0.0, 0.0], [rx, ry, 2.0], size=(n_frames, 3))
        near = np.linalg.norm(tgt - eye, axis=1) < 0.5
        tgt[near] = ctr
    c2w = np.round(A_inv[None] @ look_at_c2w_batch(eye, tgt), 6)
    c2w[:, 3] = [0, 0, 0, 1]
    if invalid_pose_every:
        c2w[invalid_pose_every - 1::invalid_pose_every] = -np.inf
    sid = scene_id or f"scene{seed:04d}_00"
    # every-5th-frame ids like the *_D5 info files (update_info_file_with_images.py)
    image_ids = [f"{5 * f:05d}" for f in range(n_frames)]
    return SynthScene(sid, pts, inst, n_objects, A, c2w, SCANNET_K.copy(), image_ids=image_ids)


@dataclass
class SynthVideo:
    video_id: str
    tracks_xyz: np.ndarray      # (T,N,3) float32 camera coordinates
    visibility: np.ndarray      # (T,N) bool
    w2c: np.ndarray             # (T,4,4) float64
    fxfycxcy: np.ndarray        # (4,) float64
    image_hw: tuple


def make_video(seed: int, n_frames: int = 150, n_tracks: int = 256, n_groups: int = 6,
               image_hw=(360, 640)) -> SynthVideo:
    """TAPVid-3D-shaped clip: rigid groups + static background, random-walk camera."""
    rng = np.random.default_rng(seed)
    H, W = image_hw
    fx = fy = 0.9 * W
    intr = np.array([fx, fy, W / 2, H / 2], dtype=np.float64)
    base = rng.uniform([-2, -1.2, 2.0], [2, 1.2, 6.0], size=(n_tracks, 3))
    group = rng.integers(0, n_groups + 1, n_tracks)          # 0 = static background
    world = np.repeat(base[None], n_frames, 0)
    for g in range(1, n_groups + 1):
        vel = rng.normal(0, 0.01, 3)
        drift = np.cumsum(rng.normal(vel, 0.004, (n_frames, 3)), 0)
        world[:, group == g] += drift[:, None, :]
    w2c = np.empty((n_frames, 4, 4))
    pos = np.zeros(3)
    yaw = 0.0
    for t in range(n_frames):
        pos = pos + rng.normal(0, 0.01, 3)
        yaw += rng.normal(0, 0.004)
        c, s = np.cos(yaw), np.sin(yaw)
        c2w = np.eye(4)
        c2w[:3, :3] = [[c, 0, s], [0, 1, 0], [-s, 0, c]]
        c2w[:3, 3] = pos
        w2c[t] = np.linalg.inv(c2w)
    hom = np.concatenate([world, np.ones((n_frames, n_tracks, 1))], 2)
    cam = np.einsum("tij,tnj->tni", w2c, hom)[..., :3].astype(np.float32)
    u = fx * cam[..., 0] / cam[..., 2] + W / 2
    v = fy * cam[..., 1] / cam[..., 2] + H / 2
    vis = (cam[..., 2] > 0.1) & (u >= 2) & (u < W - 2) & (v >= 2) & (v < H - 2)
    vis &= rng.random((n_frames, n_tracks)) > 0.1
    return SynthVideo(f"video{seed:05d}", cam, vis, w2c, intr, image_hw)


def make_stress_scene(seed: int, n_points: int, n_frames: int, room=(8.0, 8.0, 3.0)) -> SynthScene:
    """BASELINE.json configs[4] ("synthetic stress"): ``n_points`` uniform random points INSIDE the room volume (no surface
    structure, no locality in the point order) and independent inward-looking SE(3) poses; no instances.  The depth maps
    are rendered from the points themselves (msp_zbuffer_render), as in every synthetic config."""
    rng = np.random.default_rng(seed)
    rx, ry, rz = room
    pts = rng.uniform([0.0, 0.0, 0.0], [rx, ry, rz], size=(n_points, 3)).astype(np.float32)
    inst = np.zeros(n_points, np.uint16)
    ang = rng.uniform(0, 2 * np.pi)
    A = np.eye(4)
    A[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    A[:3, 3] = rng.uniform(-4, 4, 3)
    A = np.round(A, 6)
    A_inv = np.linalg.inv(A)
    eye = rng.uniform([0.5, 0.5, 0.5], [rx - 0.5, ry - 0.5, rz - 0.5], size=(n_frames, 3))
    tgt = rng.uniform([0.0, 0.0, 0.0], [rx, ry, rz], size=(n_frames, 3))
    near = np.linalg.norm(tgt - eye, axis=1) < 0.5
    tgt[near] = np.array([rx / 2, ry / 2, rz / 2])
    c2w = np.round(A_inv[None] @ look_at_c2w_batch(eye, tgt), 6)
    c2w[:, 3] = [0, 0, 0, 1]
    image_ids = [f"{5 * f:05d}" for f in range(n_frames)]
    return SynthScene(f"stress{seed:04d}_00", pts, inst, 0, A, c2w, SCANNET_K.copy(), image_ids=image_ids)
