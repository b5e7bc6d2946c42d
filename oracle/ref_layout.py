"""Write a synthetic scene to disk in the reference's ScanNet layout.

TEST INFRASTRUCTURE (used by the golden generator, the drop-in handler tests
and the CPU-baseline leg).  Layout per spatial_engine/utils/scannet_utils/README.md:104-123
and the info-dict schema in info_handler.py:7-30:

  <root>/scannet_instance_data/<scene>/aligned_points.npy   (N,6) float64 xyz+rgb
  <root>/scannet_instance_data/<scene>/instance_mask.npy    (N,)  1-indexed ids
  <root>/posed_images/<scene>/<id:05d>.jpg                  colour (only its shape is read)
  <root>/posed_images/<scene>/<id:05d>.png                  uint16 depth, millimetres
  <root>/scannet_instance_data/<name>.pkl                   {scene_id: info dict}
"""
from __future__ import annotations

import os
import pickle

import cv2
import numpy as np


def scene_info_dict(scene, root: str) -> dict:
    images_info = {}
    for f, image_id in enumerate(scene.image_ids):
        images_info[image_id] = {
            "image_path": os.path.join(root, "posed_images", scene.scene_id, f"{image_id}.jpg"),
            "depth_image_path": os.path.join(root, "posed_images", scene.scene_id, f"{image_id}.png"),
            "extrinsic_matrix": scene.c2w[f].copy(),
        }
    info = {
        "num_posed_images": scene.n_frames,
        "intrinsic_matrix": scene.K.copy(),
        "images_info": images_info,
        "axis_align_matrix": scene.axis_align.copy(),
        "num_objects": scene.num_objects,
    }
    pts = scene.points.astype(np.float64)
    for o in range(scene.num_objects):
        sel = pts[scene.instance_ids == o + 1]
        if len(sel) == 0:
            lo = hi = np.zeros(3)
        else:
            lo, hi = sel.min(0), sel.max(0)
        bbox = np.concatenate([(lo + hi) / 2, hi - lo, [float(o % 18 + 1)]])
        info[o] = {"aligned_bbox": bbox, "unaligned_bbox": bbox.copy(), "raw_category": f"thing{o}"}
    return info


def write_scene(scene, root: str, write_images: bool = True) -> None:
    inst_dir = os.path.join(root, "scannet_instance_data", scene.scene_id)
    img_dir = os.path.join(root, "posed_images", scene.scene_id)
    os.makedirs(inst_dir, exist_ok=True)
    os.makedirs(img_dir, exist_ok=True)
    n = scene.n_points
    pts6 = np.concatenate([scene.points.astype(np.float64), np.full((n, 3), 128.0)], axis=1)
    np.save(os.path.join(inst_dir, "aligned_points.npy"), pts6)
    np.save(os.path.join(inst_dir, "instance_mask.npy"), scene.instance_ids.astype(np.int64))
    if not write_images:
        return
    H, W = scene.image_hw
    blank = np.zeros((H, W, 3), np.uint8)
    ok, jpg = cv2.imencode(".jpg", blank, [cv2.IMWRITE_JPEG_QUALITY, 30])
    assert ok
    jpg = jpg.tobytes()
    for f, image_id in enumerate(scene.image_ids):
        with open(os.path.join(img_dir, f"{image_id}.jpg"), "wb") as fh:
            fh.write(jpg)
        if scene.depth is not None:
            cv2.imwrite(os.path.join(img_dir, f"{image_id}.png"), scene.depth[f])


def write_split(scenes, root: str, name: str = "scenes_synth_info_i_D5.pkl") -> str:
    infos = {}
    for s in scenes:
        write_scene(s, root)
        infos[s.scene_id] = scene_info_dict(s, root)
    path = os.path.join(root, "scannet_instance_data", name)
    with open(path, "wb") as f:
        pickle.dump(infos, f)
    return path


# ------------------------------------------------------------------------------------------------
# raw ScanNet scans (inputs of the ingestion scripts): ply + segs + aggregation + meta txt + .sens
# ------------------------------------------------------------------------------------------------
LABEL_TSV_ROWS = [("chair", 5), ("table", 7), ("wall", 1), ("floor", 2), ("sofa", 6), ("tv stand", 40), ("lamp", 35)]


def write_label_map(path: str) -> str:
    """scannetv2-labels.combined.tsv with the two columns the reference reads (SU:118-127) among others."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        f.write("id\traw_category\tcategory\tcount\tnyu40id\n")
        for k, (name, nyu) in enumerate(LABEL_TSV_ROWS):
            f.write(f"{k + 1}\t{name}\t{name}\t{10 * k}\t{nyu}\n")
    return path


def write_raw_scan(scans_dir: str, scan_name: str, seed: int = 0, n_vertices: int = 600, n_objects: int = 6,
                   n_frames: int = 7, color_hw=(24, 32), depth_hw=(12, 16), with_alignment: bool = True) -> dict:
    """One synthetic raw scan in ScanNet's release layout; returns what was written (for assertions)."""
    import json
    import struct
    import zlib
    rng = np.random.default_rng(seed)
    d = os.path.join(scans_dir, scan_name)
    os.makedirs(d, exist_ok=True)
    base = os.path.join(d, scan_name)
    # mesh: float32 xyz + uchar rgba vertices, a few triangular faces (skipped by vertex readers)
    xyz = rng.uniform(-3, 3, (n_vertices, 3)).astype(np.float32)
    rgba = rng.integers(0, 256, (n_vertices, 4)).astype(np.uint8)
    faces = rng.integers(0, n_vertices, (n_vertices // 2, 3)).astype(np.int32)
    with open(base + "_vh_clean_2.ply", "wb") as f:
        f.write((f"ply\nformat binary_little_endian 1.0\ncomment synthetic\nelement vertex {n_vertices}\n"
                 "property float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\n"
                 "property uchar blue\nproperty uchar alpha\n"
                 f"element face {len(faces)}\nproperty list uchar int vertex_indices\nend_header\n").encode())
        v = np.zeros(n_vertices, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("a", "u1")])
        v["x"], v["y"], v["z"] = xyz.T
        v["r"], v["g"], v["b"], v["a"] = rgba.T
        f.write(v.tobytes())
        for tri in faces:
            f.write(struct.pack("<B3i", 3, *tri))
    # over-segmentation: ~4 vertices per segment, non-contiguous segment ids
    n_seg = max(n_objects * 3, n_vertices // 4)
    seg_ids = np.sort(rng.choice(np.arange(5 * n_seg), n_seg, replace=False))
    seg_of_vertex = seg_ids[rng.integers(0, n_seg, n_vertices)]
    with open(base + "_vh_clean_2.0.010000.segs.json", "w") as f:
        json.dump({"params": {}, "sceneId": scan_name, "segIndices": seg_of_vertex.tolist()}, f)
    used = np.unique(seg_of_vertex)
    rng.shuffle(used)
    parts = np.array_split(used[: int(len(used) * 0.8)], n_objects)          # 20 % of the segments stay unannotated
    labels = [LABEL_TSV_ROWS[k % 5][0] for k in range(n_objects)]            # repeated labels: aliasing path of SU:30-33
    groups = [{"id": k, "objectId": k, "segments": [int(s) for s in parts[k]], "label": labels[k]} for k in range(n_objects)]
    with open(base + ".aggregation.json", "w") as f:
        json.dump({"sceneId": scan_name, "appId": "synthetic", "segGroups": groups}, f)
    A = np.eye(4)
    if with_alignment:
        th = rng.uniform(0, 2 * np.pi)
        A[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
        A[:3, 3] = rng.uniform(-2, 2, 3)
        A = np.array([[float("%f" % v) for v in row] for row in A])
    with open(base + ".txt", "w") as f:
        if with_alignment:
            f.write("axisAlignment = " + " ".join("%f" % v for v in A.ravel()) + " \n")
        f.write(f"colorHeight = {color_hw[0]}\nsceneType = Synthetic\n")
    # .sens (version 4): jpeg colour, zlib_ushort depth
    intr_c = np.array([[30.0, 0, color_hw[1] / 2, 0], [0, 30.0, color_hw[0] / 2, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    intr_d = intr_c.copy()
    poses, depths, jpgs = [], [], []
    with open(base + ".sens", "wb") as f:
        name = b"StructureSensor"
        f.write(struct.pack("<IQ", 4, len(name)) + name)
        for m in (intr_c, np.eye(4, dtype=np.float32), intr_d, np.eye(4, dtype=np.float32)):
            f.write(m.astype("<f4").tobytes())
        f.write(struct.pack("<iiIIIIfQ", 2, 1, color_hw[1], color_hw[0], depth_hw[1], depth_hw[0], 1000.0, n_frames))
        for k in range(n_frames):
            c2w = np.eye(4, dtype=np.float32)
            c2w[:3, 3] = rng.uniform(-2, 2, 3)
            th = rng.uniform(0, 2 * np.pi)
            c2w[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
            if k == 3:
                c2w[:] = -np.inf                                         # ScanNet marks lost tracking like this
            depth = rng.integers(0, 6000, depth_hw).astype("<u2")
            img = rng.integers(0, 256, (*color_hw, 3)).astype(np.uint8)
            ok, jpg = cv2.imencode(".jpg", img)
            assert ok
            zd = zlib.compress(depth.tobytes())
            f.write(c2w.astype("<f4").tobytes())
            f.write(struct.pack("<QQQQ", 1000 + k, 2000 + k, len(jpg), len(zd)))
            f.write(jpg.tobytes())
            f.write(zd)
            poses.append(c2w), depths.append(depth.astype(np.uint16)), jpgs.append(jpg.tobytes())
    return {"xyz": xyz, "rgba": rgba, "seg_of_vertex": seg_of_vertex, "groups": groups, "axis_align": A,
            "poses": np.stack(poses), "depths": np.stack(depths), "jpgs": jpgs, "intrinsic_color": intr_c}
