"""Goldens of the three L2 scripts' per-scene workers and helpers, produced by the UNMODIFIED reference
(build container only; needs /root/reference):

  make_visibility_info.process_scene            (MVI:75-125)   result digest + warning file text
  calculate_frames_relations.process_scene      (CFR:139-197)  every pair's four values + warning file text
  calculate_frames_relations.extract_yaw_pitch / calculate_camera_overlap   (CFR:86-137)
  compute_object_visibility.process_scene       (COV:72-152)   result + warnings list

on one synthetic disk scene built to reach the warning branches: two frames look away from the scene from outside the room
(no visible point: "no in bound points", NaN overlap between them: "has something wrong"), two frames have -inf poses, one
object is a wall (non-informative category) and one object has no points.

    python tests/golden/make_golden_l2_scripts.py   ->  tests/golden/l2_scripts.json.gz
"""
import contextlib
import gzip
import hashlib
import io
import json
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from multi_spatialmllm_b200 import synth  # noqa: E402
from oracle import geometry as og  # noqa: E402
from oracle import ref_layout, ref_shim  # noqa: E402

BLIND = (2, 6)


def scene():
    s = synth.make_scene(31, n_points=2500, n_frames=10, n_objects=5, scene_id="scene0031_00", invalid_pose_every=4)
    for f in BLIND:                                          # outside the room, looking further away along the view axis
        s.c2w[f, :3, 3] = np.round(s.c2w[f, :3, 3] + 60.0 * s.c2w[f, :3, 2], 6)
    s.instance_ids = np.where(s.instance_ids == 5, 0, s.instance_ids).astype(s.instance_ids.dtype)     # object 4: no points
    pts = s.points.astype(np.float64)
    H, W = s.image_hw
    depth = np.zeros((s.n_frames, *s.depth_hw), np.uint16)
    with np.errstate(all="ignore"):
        for f in range(s.n_frames):
            if np.isfinite(s.c2w[f]).all():
                depth[f] = og.depth_to_u16(og.zbuffer_render(pts, s.K, og.aligned_pose(s.axis_align, s.c2w[f]), H, W, *s.depth_hw), 1)
    s.depth = depth
    return s


def write(s, root):
    info = ref_layout.write_split([s], root, name="scenes_val_info_i_D5.pkl")
    with open(info, "rb") as f:
        infos = pickle.load(f)
    infos[s.scene_id][1]["raw_category"] = "wall"            # COV:58
    with open(info, "wb") as f:
        pickle.dump(infos, f)
    return info


def digest(obj) -> str:
    return hashlib.sha256(json.dumps(obj, sort_keys=True).encode()).hexdigest()


def main():
    ref = ref_shim.import_reference()
    quiet = contextlib.redirect_stdout(io.StringIO())
    with quiet:
        from spatial_engine.object_perception import compute_object_visibility as cov
    s = scene()
    out = {}
    with tempfile.TemporaryDirectory() as tmp, np.errstate(all="ignore"):
        root = os.path.join(tmp, "data", "scannet")
        info = write(s, root)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            with quiet:
                h = ref.info_handler.SceneInfoHandler(info)
                w1, w2 = os.path.join(tmp, "mvi_warn.txt"), os.path.join(tmp, "cfr_warn.txt")
                _, vis = ref.mvi.process_scene(s.scene_id, h, w1)
                _, rel = ref.cfr.process_scene(s.scene_id, h, w2)
                vdict = {f"{s.scene_id}:image_to_points:{k}": json.dumps(v) for k, v in vis["image_to_points"].items()}
                _, ov, ov_warn = cov.process_scene(s.scene_id, h, vdict)
            ids = h.get_all_extrinsic_valid_image_ids(s.scene_id)
            masks = {i: np.isin(np.arange(s.n_points), vis["image_to_points"][i]) for i in ids}
            out["visibility"] = {"image_to_points_sha256": digest(vis["image_to_points"]),
                                 "point_to_images_sha256": digest({str(k): v for k, v in vis["point_to_images"].items()}),
                                 "visible_counts": {i: len(v) for i, v in vis["image_to_points"].items()},
                                 "warnings": open(w1).read() if os.path.exists(w1) else ""}
            out["relations"] = {"pairs": [[a, b, v["overlap"], v["distance"], v["yaw"], v["pitch"]] for (a, b), v in rel.items()],
                                "warnings": open(w2).read() if os.path.exists(w2) else ""}
            out["yaw_pitch"] = [[i, *[float(x) for x in ref.cfr.extract_yaw_pitch(h.get_extrinsic_matrix_align(s.scene_id, i))]] for i in ids]
            out["overlap_calls"] = [[a, b, float(ref.cfr.calculate_camera_overlap(masks, a, b))] for a, b in
                                    ((ids[0], ids[1]), (ids[1], ids[4]), ("00010", ids[0]), ("00010", "00030"))
                                    if a in masks and b in masks]
            out["object_visibility"] = {"result": json.loads(json.dumps(ov)), "warnings": ov_warn}
            out["valid_ids"] = ids
        finally:
            os.chdir(cwd)
    print({k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})
    print("MVI warnings:", out["visibility"]["warnings"], "CFR warnings:", out["relations"]["warnings"][:400], out["object_visibility"]["warnings"])
    with gzip.open(os.path.join(HERE, "l2_scripts.json.gz"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
