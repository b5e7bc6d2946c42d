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
