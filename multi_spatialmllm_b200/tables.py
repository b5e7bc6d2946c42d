"""L2 tables in the reference's on-disk schemas + the per-split drivers that fill them.

  visibility table  parquet, columns key / values; key "{scene}:image_to_points:{image_id}" or
                    "{scene}:point_to_images:{point}", values = JSON string (make_visibility_info.py:59-65 —
                    the canonical form every reader expects, info_handler.py:510-513; the keys written by
                    run_split at make_visibility_info.py:163-172 are not readable by the reference itself)
  overlap table     parquet, columns scene_id,image_id1,image_id2,overlap,distance,yaw,pitch
                    (calculate_frames_relations.py:43-51) + the "_nonzero" companion (:59-84)
  object visibility pkl {scene: {"object_to_images": {...}, "image_to_objects": {...}}}
                    (compute_object_visibility.py:14-38,139-150)

Geometry comes from engine.DeviceScene (CUDA); this module is host-side formatting and file I/O only.
"""
from __future__ import annotations

import json
import os
import pickle

import numpy as np
import pandas as pd

from . import engine

NONINFORMATIVE_DESC = {"wall", "object", "floor", "ceiling", "window"}     # compute_object_visibility.py:58


def visibility_rows(scene_id: str, table: dict):
    for image_id, pts in table["image_to_points"].items():
        yield f"{scene_id}:image_to_points:{image_id}", json.dumps(pts)
    for point, images in table["point_to_images"].items():
        yield f"{scene_id}:point_to_images:{point}", json.dumps(images)


def save_visibility_parquet(scene_tables: dict, parquet_file: str) -> int:
    rows = [r for sid, t in scene_tables.items() for r in visibility_rows(sid, t)]
    os.makedirs(os.path.dirname(parquet_file) or ".", exist_ok=True)
    pd.DataFrame(rows, columns=["key", "values"]).to_parquet(parquet_file, index=False)
    return len(rows)


def overlap_dataframe(overlap_info: dict, nonzero_only: bool = False) -> pd.DataFrame:
    cols = {k: [] for k in ("scene_id", "image_id1", "image_id2", "overlap", "distance", "yaw", "pitch")}
    for sid, pairs in overlap_info.items():
        for (a, b), v in pairs.items():
            if nonzero_only and not (v["overlap"] != 0.0):      # NaN != 0.0 is True: kept, like CFR:67
                continue
            cols["scene_id"].append(sid), cols["image_id1"].append(a), cols["image_id2"].append(b)
            for k in ("overlap", "distance", "yaw", "pitch"):
                cols[k].append(v[k])
    return pd.DataFrame(cols)


def save_overlap_parquet(overlap_info: dict, parquet_file: str):
    """Writes <file> and <file>_nonzero.parquet like run_split (calculate_frames_relations.py:241-245)."""
    os.makedirs(os.path.dirname(parquet_file) or ".", exist_ok=True)
    df = overlap_dataframe(overlap_info)
    if len(df):
        df.to_parquet(parquet_file, index=False)
    nz = overlap_dataframe(overlap_info, nonzero_only=True)
    if len(nz):
        nz.to_parquet(parquet_file.replace(".parquet", "_nonzero.parquet"), index=False)
    return len(df), len(nz)


def device_scene_from_handler(scene_infos, scene_id: str, device="cuda:0"):
    """Upload one scene of a SceneInfoHandler: aligned points, every frame's pose (invalid ones included, they
    are masked on the device) and its depth PNG.  Returns (DeviceScene, image_ids in row order)."""
    image_ids = scene_infos.get_all_image_ids(scene_id)
    pts = np.ascontiguousarray(scene_infos.get_scene_points_align(scene_id)[:, :3], dtype=np.float64)
    c2w = np.stack([scene_infos.get_extrinsic_matrix(scene_id, i, warning=False) for i in image_ids])
    H, W = scene_infos.get_image_shape(scene_id)
    depth = None
    for f, i in enumerate(image_ids):
        d = scene_infos.get_depth_image(scene_id, i)
        if d is None:
            continue
        if depth is None:
            depth = np.zeros((len(image_ids), *d.shape[:2]), np.uint16)
        depth[f] = d
    inst = n_obj = None
    try:
        inst = scene_infos.get_scene_instance_mask(scene_id).astype(np.uint16)
        n_obj = int(scene_infos.get_num_objects(scene_id))
    except Exception:
        pass
    sc = engine.DeviceScene(pts, scene_infos.get_intrinsic_matrix(scene_id), c2w,
                            axis_align=scene_infos.get_world_to_axis_align_matrix(scene_id), depth=depth,
                            instance_ids=inst, num_objects=n_obj or 0, image_hw=(H, W), depth_hw=depth.shape[1:],
                            depth_scale=scene_infos.depth_value_scale, device=device)
    return sc, image_ids


def run_visibility_split(scene_infos, output_file: str, scene_ids=None, device="cuda:0") -> dict:
    """make_visibility_info.run_split (MVI:127-177) on the GPU; writes the canonical parquet."""
    tables = {}
    for sid in (scene_ids if scene_ids is not None else scene_infos.get_all_scene_ids()):
        sc, image_ids = device_scene_from_handler(scene_infos, sid, device)
        tables[sid] = engine.process_scene_visibility(sc, image_ids)
    save_visibility_parquet(tables, output_file)
    return tables


def run_relations_split(scene_infos, output_parquet: str, scene_ids=None, device="cuda:0") -> dict:
    """calculate_frames_relations.run_split (CFR:200-253) on the GPU."""
    info = {}
    for sid in (scene_ids if scene_ids is not None else scene_infos.get_all_scene_ids()):
        sc, image_ids = device_scene_from_handler(scene_infos, sid, device)
        info[sid] = engine.process_scene_relations(sc, image_ids)
    save_overlap_parquet(info, output_parquet)
    return info


def object_visibility_scene(scene_infos, scene_id: str, sc: "engine.DeviceScene", image_ids) -> dict:
    """compute_object_visibility.process_scene (COV:72-152) from the device's per-(frame, object) counts."""
    result = {"object_to_images": {}, "image_to_objects": {}}
    sc.project_visibility()
    counts = sc.object_counts().cpu().numpy()
    valid = sc.valid.cpu().numpy().astype(bool)
    inst = scene_infos.get_scene_instance_mask(scene_id)
    for o in range(scene_infos.get_num_objects(scene_id)):
        if scene_infos.get_object_raw_category(scene_id, o) in NONINFORMATIVE_DESC:
            continue
        total = int(np.count_nonzero(inst == o + 1))
        if total == 0:
            continue
        thr = max(1, int(0.05 * total))
        for f, image_id in enumerate(image_ids):
            c = int(counts[f, o])
            if valid[f] and c >= thr:
                row = {"intersection_count": c, "visibility": (c / total) * 100.0}
                result["object_to_images"].setdefault(o, []).append({"image_id": image_id, **row})
                result["image_to_objects"].setdefault(image_id, []).append({"object_id": o, **row})
    return result


def run_object_visibility_split(scene_infos, output_pkl: str, scene_ids=None, device="cuda:0") -> dict:
    out = {}
    for sid in (scene_ids if scene_ids is not None else scene_infos.get_all_scene_ids()):
        sc, image_ids = device_scene_from_handler(scene_infos, sid, device)
        out[sid] = object_visibility_scene(scene_infos, sid, sc, image_ids)
    os.makedirs(os.path.dirname(output_pkl) or ".", exist_ok=True)
    with open(output_pkl, "wb") as f:
        pickle.dump(out, f)
    return out


COVERAGE_TOLERANCE = 0.1        # single_object_coverage_finder.py TOLERANCE


def object_coverage_scene(scene_infos, scene_id: str, sc: "engine.DeviceScene", image_ids, object_visibility: dict,
                          max_images: int = 5) -> dict:
    """single_object_coverage_finder.process_scene_for_coverage (SOCF:267-284) with process_object (SOCF:222-265):
    per object {"height" | "length" | "width": {k: [image-id tuples]}}, using one msp_object_extents pass for the scene
    instead of union masks.  Uses the global ``random`` state like the reference (seed it the same way)."""
    from . import coverage
    sc.project_visibility()
    ext = sc.object_extents().cpu().numpy()                      # (F, n_obj, 6)
    row = {img: f for f, img in enumerate(image_ids)}
    out = {}
    for object_id, vis_list in object_visibility["object_to_images"].items():
        imgs = [v["image_id"] for v in vis_list]
        rows = [row[i] for i in imgs]
        width_axis = scene_infos.get_object_width_axis_aligned(scene_id, object_id)
        plan = (("height", 2, scene_infos.get_object_height(scene_id, object_id)),
                ("length", 1 if width_axis == 0 else 0, scene_infos.get_object_length(scene_id, object_id)),
                ("width", width_axis, scene_infos.get_object_width(scene_id, object_id)))
        res = {}
        for name, axis, target in plan:
            res[name] = coverage.find_minimal_combinations(imgs, ext[rows, object_id, axis], ext[rows, object_id, 3 + axis],
                                                           float(target), COVERAGE_TOLERANCE, max_images=max_images)
        out[object_id] = res
    return out
