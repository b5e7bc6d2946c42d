"""CPU oracle for the MultiSPA geometry hot path (NumPy, float64).

TEST INFRASTRUCTURE ONLY.  Nothing under ``multi_spatialmllm_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only
as the checker or the timed CPU baseline, never as the product path.

Every function restates one piece of the reference's algorithm and cites the
reference file:line it follows (paths relative to the reference checkout):

  IH  = spatial_engine/utils/scannet_utils/handler/info_handler.py
  CFR = spatial_engine/camera_movement/calculate_frames_relations.py
  CME = spatial_engine/camera_movement/camera_movement_engine_train_val.py
  MVI = spatial_engine/utils/scannet_utils/make_visibility_info.py
  VC  = spatial_engine/visual_correspondence/visual_correspondence_qa_engine_coor_2_coor.py
  DE  = spatial_engine/depth_perception/depth_estimation_dot_engine.py
  DC  = spatial_engine/depth_perception/depth_comparison_coor_engine.py
  COV = spatial_engine/object_perception/compute_object_visibility.py
  OM  = spatial_engine/object_movement/single_object_movement_engine_coord.py

Parity status: PINNED for every function except ``zbuffer_render`` /
``depth_to_u16``.  The reference ships no golden vectors or tests (SURVEY.md
§4), so the pin is made by ``tests/golden/make_golden.py``, which imports the
unmodified reference from ``/root/reference`` through ``oracle/ref_shim.py``,
runs it on seeded synthetic scenes written in the reference's on-disk layout
and stores its outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this module against those files (and against the live reference when
``/root/reference`` is present).  ``zbuffer_render`` has no counterpart in the
reference: its parity is UNPINNED and it is only an oracle for
our own kernel's definition.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# projection  (IH:46-72, IH:313-335)
# ----------------------------------------------------------------------------

def project_points(points_h: np.ndarray, K: np.ndarray, E_c2w: np.ndarray):
    """World -> image.  Follows IH:46-72 step by step.

    points_h (N,4) homogeneous float64, K (4,4), E_c2w (4,4) camera->world.
    Returns xy (N,2) float64 in colour-pixel units and z (N,) float64 metres.
    The order of operations matters for bit parity of the indices derived
    from xy: invert E (IH:57), cam = Einv @ P^T (IH:60), z = cam[2] (IH:63),
    img = K @ cam (IH:66), img /= img[2] (IH:69).
    """
    w2c = np.linalg.inv(E_c2w)
    cam = w2c @ points_h.T
    z = cam[2, :]
    img = K @ cam
    img = img / img[2, :]
    return img.T[:, :2], z


def homogenize(points_xyz: np.ndarray) -> np.ndarray:
    """(3,) or (N,3) -> (N,4) with a ones column (IH:328-331)."""
    p = np.asarray(points_xyz)
    if p.ndim == 1:
        p = p[None, :]
    return np.hstack([p, np.ones((p.shape[0], 1))])


def project_3d_point_to_image(points_xyz, K, E_c2w):
    """IH:313-335 with the pose already selected (aligned or raw)."""
    return project_points(homogenize(points_xyz), K, E_c2w)


def aligned_pose(axis_align: np.ndarray, c2w: np.ndarray) -> np.ndarray:
    """IH:113-124: aligned camera->world = axis_align @ c2w."""
    return axis_align @ c2w


def pose_is_valid(c2w: np.ndarray) -> bool:
    """IH:409-418: a frame is dropped when its pose holds inf or nan."""
    return not (np.any(np.isinf(c2w)) or np.any(np.isnan(c2w)))


# ----------------------------------------------------------------------------
# visibility  (IH:337-386)
# ----------------------------------------------------------------------------

def in_image_boundary(xy: np.ndarray, image_h: int, image_w: int) -> np.ndarray:
    """IH:337-344: 0 <= u < W and 0 <= v < H on the float coordinates."""
    u, v = xy[:, 0], xy[:, 1]
    return (u >= 0) & (u < image_w) & (v >= 0) & (v < image_h)


def depth_pixel_index(xy, image_h, image_w, depth_h, depth_w):
    """IH:359-366: scale to depth resolution, round half-even, cast, clip."""
    sx = depth_w / image_w
    sy = depth_h / image_h
    with np.errstate(invalid="ignore", over="ignore"):
        dx = np.round(xy[:, 0] * sx).astype(int)
        dy = np.round(xy[:, 1] * sy).astype(int)
    dx = np.clip(dx, 0, depth_w - 1)
    dy = np.clip(dy, 0, depth_h - 1)
    return dy, dx


def visible_by_depth(xy, z, depth_u16, image_h, image_w, depth_scale=0.001):
    """IH:346-373: 0 < z < D[dy,dx]*scale  (strict on both sides)."""
    dh, dw = depth_u16.shape[:2]
    dy, dx = depth_pixel_index(xy, image_h, image_w, dh, dw)
    d = depth_u16[dy, dx] * depth_scale
    return (z > 0) & (z < d)


def check_point_visibility(xy, z, depth_u16, image_h, image_w, depth_scale=0.001):
    """IH:375-386: in-bounds AND depth test."""
    return in_image_boundary(xy, image_h, image_w) & visible_by_depth(
        xy, z, depth_u16, image_h, image_w, depth_scale)


def frame_visibility(points_xyz, K, E_c2w, depth_u16, image_h, image_w, depth_scale=0.001):
    """One iteration of the MVI:93-100 / CFR:152-157 frame loop."""
    with np.errstate(divide="ignore", invalid="ignore"):
        xy, z = project_3d_point_to_image(points_xyz, K, E_c2w)
        mask = check_point_visibility(xy, z, depth_u16, image_h, image_w, depth_scale)
    return xy, z, mask


def pack_mask_bits(mask: np.ndarray, words: int | None = None) -> np.ndarray:
    """bool (N,) -> uint32 words, bit p%32 of word p//32 = mask[p] (our layout)."""
    n = mask.shape[0]
    nw = (n + 31) // 32 if words is None else words
    padded = np.zeros(nw * 32, dtype=np.uint8)
    padded[:n] = mask
    return np.packbits(padded.reshape(nw, 32), axis=1, bitorder="little").view("<u4").reshape(nw)


def unpack_mask_bits(words: np.ndarray, n: int) -> np.ndarray:
    b = np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")
    return b[:n].astype(bool)


# ----------------------------------------------------------------------------
# per-frame pose scalars and per-pair relations  (CFR:86-137, 176-189)
# ----------------------------------------------------------------------------

def extract_yaw_pitch(E: np.ndarray):
    """CFR:86-100: viewing direction = third column of the rotation block."""
    R = E[:3, :3] if E.shape == (4, 4) else E
    zaxis = R[:, 2]
    yaw = np.degrees(np.arctan2(zaxis[1], zaxis[0]))
    pitch = np.degrees(np.arcsin(zaxis[2] / np.linalg.norm(zaxis)))
    return yaw, pitch


def camera_overlap(mask_a: np.ndarray, mask_b: np.ndarray) -> float:
    """CFR:128-137: |A&B| / |A|B| * 100; NaN when the union is empty."""
    inter = np.sum(np.logical_and(mask_a, mask_b))
    uni = np.sum(np.logical_or(mask_a, mask_b))
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / uni * 100


def pair_relations(E1, E2, mask1, mask2):
    """CFR:176-189 for one ordered pair (frame1 listed before frame2)."""
    y1, p1 = extract_yaw_pitch(E1)
    y2, p2 = extract_yaw_pitch(E2)
    return {
        "overlap": camera_overlap(mask1, mask2),
        "distance": np.linalg.norm(E2[:3, 3] - E1[:3, 3]),
        "yaw": y2 - y1,
        "pitch": p2 - p1,
    }


def relative_displacement(E1, E2):
    """CME:177-191: translation of inv(E1) @ E2 (second camera in first's frame)."""
    rel = np.linalg.inv(E1) @ E2
    d = rel[:3, 3]
    return d, np.linalg.norm(d)


def camera_movement_values(E1, E2, yaw_angle, pitch_angle):
    """CME:168-172, 205-225: wrap yaw, truncate to ints the way the writer does."""
    if abs(yaw_angle) > 180:
        yaw_angle = yaw_angle - 360 if yaw_angle > 0 else yaw_angle + 360
    d, dist = relative_displacement(E1, E2)
    return {
        "x_distance": int(abs(d[0]) * 1000), "y_distance": int(abs(d[1]) * 1000),
        "z_distance": int(abs(d[2]) * 1000),
        "x_value": int(d[0] * 1000), "y_value": int(d[1] * 1000), "z_value": int(d[2] * 1000),
        "yaw_angle": int(abs(yaw_angle)), "pitch_angle": int(abs(pitch_angle)),
        "total_distance": int(dist * 1000),
        "displacement_vector": d.tolist(),
    }


# ----------------------------------------------------------------------------
# sample integers  (VC:303-344, DE:216-218,261, DC:267-276)
# ----------------------------------------------------------------------------

def common_visible_points(points1, points2) -> np.ndarray:
    """VC:303: sorted unique ids visible in both frames."""
    return np.intersect1d(points1, points2)


def permille(coord: float, extent: int) -> int:
    """VC:341-344 / DE:216-217: Python round (half-even) of coord/extent*1000."""
    return round((coord / extent) * 1000)


def depth_mm(z: float) -> int:
    """DE:218: round(z * 1000)."""
    return round(z * 1000)


def annotate_point(point_xyz, K, E_c2w, depth_u16, image_h, image_w, depth_scale=0.001):
    """IH:291-305 (check_visible=True, return_depth=True) + DE:216-218,261.

    Returns None when the re-check says invisible (engines then skip the point).
    """
    xy, z, vis = frame_visibility(np.asarray(point_xyz)[:3], K, E_c2w, depth_u16,
                                  image_h, image_w, depth_scale)
    if not vis[0]:
        return None
    u, v, zz = float(xy[0, 0]), float(xy[0, 1]), float(z[0])
    return {
        "u": u, "v": v, "z": zz,
        "x_permille": permille(u, image_w), "y_permille": permille(v, image_h),
        "depth_mm": depth_mm(zz), "pix": (int(u), int(v)),
    }


# ----------------------------------------------------------------------------
# object visibility counts  (COV:72-152, IH:250-261)
# ----------------------------------------------------------------------------

def object_visible_counts(mask: np.ndarray, instance_ids: np.ndarray, num_objects: int) -> np.ndarray:
    """COV:133-134 for all objects of one frame: |visible ∩ {p: inst[p]==o+1}|.

    instance_ids are 1-indexed with 0 = no object (IH:255).
    Returns counts (num_objects,) int64, counts[o] for object id o.
    """
    ids = instance_ids[mask].astype(np.int64)
    c = np.bincount(ids, minlength=num_objects + 1)
    return c[1:num_objects + 1]


def object_visibility_rows(counts: np.ndarray, totals: np.ndarray):
    """COV:117,134-136: keep when cnt >= max(1,int(0.05*total)); pct = cnt/total*100."""
    rows = []
    for o, (c, t) in enumerate(zip(counts.tolist(), totals.tolist())):
        if t == 0:
            continue
        thr = max(1, int(0.05 * t))
        if c >= thr:
            rows.append((o, c, (c / t) * 100.0))
    return rows


# ----------------------------------------------------------------------------
# z-buffer depth render — NOT IN THE REFERENCE (parity unpinned; our definition)
# ----------------------------------------------------------------------------

def zbuffer_render(points_xyz, K, E_c2w, image_h, image_w, depth_h, depth_w) -> np.ndarray:
    """Per-pixel nearest depth at depth resolution, float32, +inf where empty.

    A point contributes iff z > 0 and it is inside the colour image
    (IH:342-343); it lands on the pixel given by the reference's own index rule
    (IH:359-366).  The stored value is float32(z) rounded to nearest.
    """
    with np.errstate(divide="ignore", invalid="ignore"):
        xy, z = project_3d_point_to_image(points_xyz, K, E_c2w)
        ok = in_image_boundary(xy, image_h, image_w) & (z > 0)
    dy, dx = depth_pixel_index(xy[ok], image_h, image_w, depth_h, depth_w)
    buf = np.full((depth_h, depth_w), np.inf, dtype=np.float32)
    np.minimum.at(buf, (dy, dx), z[ok].astype(np.float32))
    return buf


def depth_to_u16(zbuf: np.ndarray, margin_mm: int = 1, depth_scale: float = 0.001) -> np.ndarray:
    """float32 z-buffer -> ScanNet-style uint16 millimetres, 0 = no reading.

    d_mm = floor(z / depth_scale) + 1 + margin_mm (computed in float64), clamped
    to 65535, so that the strict test z < d_mm*depth_scale (IH:371) accepts the
    surface that produced the pixel.
    """
    z = zbuf.astype(np.float64)
    filled = np.isfinite(z)
    mm = np.floor(np.where(filled, z, 0.0) / depth_scale) + 1 + margin_mm
    mm = np.minimum(mm, 65535.0)
    return np.where(filled, mm, 0).astype(np.uint16)


# ----------------------------------------------------------------------------
# TAPVid-3D tracks  (OM:293-315, 446-451, 484-498, 329-376)
# ----------------------------------------------------------------------------

def tracks_to_world(tracks_xyz_cam: np.ndarray, w2c: np.ndarray) -> np.ndarray:
    """OM:446-451: per-frame c2w = inv(w2c); world = c2w @ [xyz,1]."""
    T, N, _ = tracks_xyz_cam.shape
    c2w = np.linalg.inv(w2c)
    hom = np.concatenate([tracks_xyz_cam, np.ones((T, N, 1))], axis=2)
    world = np.einsum("nij,nkj->nki", c2w, hom)
    return world[..., :3]


def track_project_point(p_cam, fxfycxcy, image_h, image_w):
    """OM:293-315: pinhole with the 1e-8 guard; None unless 0<=u_n<1, 0<=v_n<1, z>0."""
    fx, fy, cx, cy = fxfycxcy
    # float64 throughout: under the reference's pinned numpy 1.26.4 a float32 track coordinate
    # plus the Python float 1e-8 promotes to float64 (value-based casting); NEP 50 would not.
    x, y, z = np.asarray(p_cam, dtype=np.float64)
    u = (fx * x / (z + 1e-8)) + cx
    v = (fy * y / (z + 1e-8)) + cy
    un = u / image_w
    vn = v / image_h
    if not (0 <= un < 1 and 0 <= vn < 1 and z > 0):
        return None
    return [un, vn]


def track_pair_distances(world_track: np.ndarray, visible_frames: np.ndarray):
    """OM:484-498: all i<j over a track's visible frames, ||w_j - w_i||."""
    V = len(visible_frames)
    ii, jj = np.triu_indices(V, k=1)
    f1 = visible_frames[ii]
    f2 = visible_frames[jj]
    d = np.linalg.norm(world_track[f2] - world_track[f1], axis=1)
    return d, f1, f2


def track_pair_sample_values(world, cam, w2c, fxfycxcy, f1, f2, p, image_h, image_w,
                             object_thr=0.01, camera_thr=0.01):
    """OM:329-376 for one (frame1, frame2, point) triple; None if a projection is invalid."""
    dvec = world[f2, p] - world[f1, p]
    dist = np.linalg.norm(dvec)
    moving = True
    if dist < object_thr:
        moving, dist, dvec = False, 0, np.zeros(3)
    c1 = np.linalg.inv(w2c[f1])
    c2 = np.linalg.inv(w2c[f2])
    cam_moving = bool(np.linalg.norm(c2[:3, 3] - c1[:3, 3]) >= camera_thr)
    dcam = (w2c[f1] @ np.concatenate([dvec, [0]]))[:3]
    p1 = track_project_point(cam[f1, p], fxfycxcy, image_h, image_w)
    p2 = track_project_point(cam[f2, p], fxfycxcy, image_h, image_w)
    if p1 is None or p2 is None:
        return None
    return {
        "p1": (round(p1[0] * 1000), round(p1[1] * 1000)),
        "p2": (round(p2[0] * 1000), round(p2[1] * 1000)),
        "total_distance_mm": round(dist * 1000),
        "xyz_mm": tuple(round(float(c) * 1000) for c in dcam),
        "dcam": dcam, "point_moving": int(moving), "cam_moving": int(cam_moving),
    }


def rigid_body_groups(points: np.ndarray, threshold: float = 0.1, smoothing_factor: float = 0.01, min_size: int = 5):
    """OM:34-105: accumulate |d_t - d_{t-1}| of all pair distances where it exceeds ``smoothing_factor``, average-linkage
    clustering of the accumulated matrix cut at ``threshold``, groups with more than ``min_size`` members.
    Returns (groups, condensed accumulated loss)."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import pdist
    points = np.asarray(points, dtype=np.float64)
    loss = np.zeros(points.shape[1] * (points.shape[1] - 1) // 2)
    prev = pdist(points[0])
    for t in range(1, points.shape[0]):
        cur = pdist(points[t])
        change = np.abs(cur - prev)
        loss += np.where(change > smoothing_factor, change, 0)
        prev = cur
    labels = fcluster(linkage(loss, method="average"), threshold, criterion="distance")
    groups = [np.where(labels == i)[0].tolist() for i in range(1, int(labels.max()) + 1)]
    return [g for g in groups if len(g) > min_size], loss
