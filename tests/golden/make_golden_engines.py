"""Engine-level goldens: run the reference's own QA engines (numeric core only) on the scannet_small fixture.

    python tests/golden/make_golden_engines.py        (build container only; needs /root/reference)

Writes tests/golden/engines_small.json with the NUMERIC fields of the samples the unmodified reference engines
produce (no template text is stored):
  depth_estimation_coor   DepthEstimationCoorQAEngine.generate_qa_training_single_scene   (DE-coor:175-254)
  depth_comparison_coor   DepthComparisonCoorQAEngine.generate_qa_training_single_scene   (DC-coor:240-346)
  visual_correspondence   build_training_sample                                            (VC-coor:264-394)
  camera_movement         build_training_sample for three question types                   (CME:153-245)
One shim is needed beyond mmengine: the reference's visual-correspondence engine calls
SceneInfoHandler.get_image_size, which the reference handler does not define; it is aliased to
get_image_shape on the instance.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import pickle
import random
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from multi_spatialmllm_b200 import synth  # noqa: E402
from oracle import ref_layout, ref_shim  # noqa: E402


def main():
    ref = ref_shim.import_reference()
    g = dict(np.load(os.path.join(HERE, "scannet_small.npz")))
    s = synth.SynthScene("scene0011_00", g["points"], g["instance_ids"], int(g["num_objects"]), g["axis_align"], g["c2w"],
                         g["K"], tuple(int(v) for v in g["image_hw"]), tuple(int(v) for v in g["depth_hw"]),
                         depth=g["depth"], image_ids=[str(v) for v in g["image_ids"]])
    out = {}
    quiet = contextlib.redirect_stdout(io.StringIO())
    with tempfile.TemporaryDirectory() as tmp, np.errstate(all="ignore"):
        info_path = ref_layout.write_split([s], tmp)
        with quiet:
            h = ref.info_handler.SceneInfoHandler(info_path, posed_images_root=os.path.join(tmp, "posed_images"),
                                                  instance_data_root=os.path.join(tmp, "scannet_instance_data"))
            _, vis = ref.mvi.process_scene(s.scene_id, h, os.path.join(tmp, "w.txt"))
            _, rel = ref.cfr.process_scene(s.scene_id, h, os.path.join(tmp, "w.txt"))
        pkl = os.path.join(tmp, "vis.pkl")
        with open(pkl, "wb") as f:
            pickle.dump({s.scene_id: vis}, f)
        with quiet:
            ref.mvi.convert_pkl_to_parquet(pkl)
            vih = ref.info_handler.VisibilityInfoHandler(pkl.replace(".pkl", ".parquet"))
            from spatial_engine.depth_perception import depth_estimation_coor_engine as de
            from spatial_engine.depth_perception import depth_comparison_coor_engine as dc
            from spatial_engine.visual_correspondence import visual_correspondence_qa_engine_coor_2_coor as vc
            # the camera-movement engine does `from TEMPLATES import ...` (it is run from its own directory)
            sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "spatial_engine", "camera_movement"))
            from spatial_engine.camera_movement import camera_movement_engine_train_val as cme

        # ---- depth estimation (coor)
        eng = de.DepthEstimationCoorQAEngine.__new__(de.DepthEstimationCoorQAEngine)
        eng.scene_info, eng.visibility_info = h, vih
        eng.max_samples, eng.max_n_points_per_image, eng.warning_file = -1, 3, os.path.join(tmp, "w.txt")
        eng.templates = {"questions": ["{x1} {y1}"], "answers": ["{x1} {y1} {depth}"]}
        eng.task_description = ["t"]
        random.seed(4)
        with quiet:
            samples = eng.generate_qa_training_single_scene(s.scene_id)
        out["depth_estimation_coor"] = [
            {"image_id": x["image"][0].split("/")[1][:5], "point": int(x["id"].split("point")[1]),
             "xy_permille": [int(v) for v in x["conversations"][0]["value"].split("\n")[1].split()],
             "gt_value": int(x["gt_value"]), "ori_coordinates": [int(v) for v in x["ori_coordinates"]]} for x in samples]

        # ---- depth comparison (coor)
        eng2 = dc.DepthComparisonCoorQAEngine.__new__(dc.DepthComparisonCoorQAEngine)
        eng2.scene_info, eng2.visibility_info = h, vih
        eng2.max_samples, eng2.max_n_points_per_image, eng2.warning_file = -1, 2, os.path.join(tmp, "w.txt")
        one = ["{x1} {y1} {x2} {y2}"]
        eng2.templates = {"closer_questions": one, "farther_questions": one, "closer_answers": ["{correct_x} {correct_y}"],
                          "farther_answers": ["{correct_x} {correct_y}"]}
        eng2.task_description = ["t"]
        random.seed(6)
        with quiet:
            samples = eng2.generate_qa_training_single_scene(s.scene_id)
        out["depth_comparison_coor"] = [
            {"image_id": x["image"][0].split("/")[1][:5], "points": [int(t[1:]) for t in x["id"].split("_")[-2:]],
             "gt_value": [int(v) for v in x["gt_value"]], "is_closer_question": bool(x["is_closer_question"]),
             "points_info": [{"x": int(q["x"]), "y": int(q["y"]), "depth": int(q["depth"]), "coords": [int(c) for c in q["coords"]],
                              "letter": q["letter"]} for q in x["points_info"]]} for x in samples]

        # ---- visual correspondence (coor_2_coor)
        h.get_image_size = h.get_image_shape            # reference defect, SURVEY.md §2.3
        rows = [{"scene_id": s.scene_id, "image_id1": a, "image_id2": b, **v} for (a, b), v in rel.items() if v["overlap"] > 0]
        vis_dict = {s.scene_id: vis}
        vc.TASK_DESCRIPTION, vc.TEMPLATES = ["t"], {"questions": ["q"], "answers": ["a"]}
        random.seed(1)
        res = []
        for i, row in enumerate(rows):
            with quiet:
                smp = vc.build_training_sample(h, row, i, vis_dict, os.path.join(tmp, "w.txt"), max_points_per_pair=1)
            if smp is None:
                continue
            res.append({"images": [p.split("/")[1][:5] for p in smp["image"]], "p1": [int(v) for v in smp["p1_list"][0]],
                        "p2": [int(v) for v in smp["p2_list"][0]], "gt_value": [int(v) for v in smp["gt_value"]]})
        out["visual_correspondence_coor_2_coor"] = res

        # ---- camera movement
        cme.TASK_DESCRIPTION = ["t"]
        for qt in list(cme.QUESTION_TEMPLATES):
            cme.QUESTION_TEMPLATES[qt], cme.ANSWER_TEMPLATES[qt] = ["q"], ["a"]
        random.seed(0)
        res = []
        for i, row in enumerate(rows):
            for qt in ("total_distance", "yaw_angle", "displacement_vector"):
                with quiet:
                    smp = cme.build_training_sample(h, pd.Series(row), i, qt)
                av = smp["answer_values"]
                res.append({"images": [p.split("/")[1][:5] for p in smp["image"]], "question_type": qt,
                            "row": {k: float(row[k]) for k in ("overlap", "distance", "yaw", "pitch")},
                            "answer_values": {k: (v if isinstance(v, (str, list)) else int(v)) for k, v in av.items()},
                            "gt_value": smp["gt_value"] if isinstance(smp["gt_value"], (str, list)) else int(smp["gt_value"])})
        out["camera_movement"] = res
    # ---- object coverage search (SOCF:76-220) for every object / axis of the fixture scene
    with quiet:
        from spatial_engine.object_perception import single_object_coverage_finder as socf
    with tempfile.TemporaryDirectory() as tmp2, np.errstate(all="ignore"):
        info_path = ref_layout.write_split([s], tmp2)
        with quiet:
            h2 = ref.info_handler.SceneInfoHandler(info_path, posed_images_root=os.path.join(tmp2, "posed_images"),
                                                   instance_data_root=os.path.join(tmp2, "scannet_instance_data"))
            _, vis2 = ref.mvi.process_scene(s.scene_id, h2, os.path.join(tmp2, "w.txt"))
        vdict = {f"{s.scene_id}:image_to_points:{k}": json.dumps(v) for k, v in vis2["image_to_points"].items()}
        scene_pts = h2.get_scene_points_align(s.scene_id)[:, :3]
        cov = []
        for obj in range(s.num_objects):
            idx = h2.get_object_point_index(s.scene_id, obj)
            imgs = [k for k, v in vis2["image_to_points"].items() if len(set(v) & set(idx.tolist())) > 0]
            bbox = h2.get_object_gt_bbox(s.scene_id, obj)
            for axis in (0, 1, 2):
                for tol in (0.1, 0.35):
                    random.seed(0)
                    with quiet:
                        sol = socf.find_minimal_combinations(s.scene_id, scene_pts, idx, imgs, vdict, axis, float(bbox[3 + axis]), tol, max_images=5)
                    cov.append({"object": obj, "axis": axis, "target": float(bbox[3 + axis]), "tolerance": tol, "images": imgs,
                                "solutions": {str(k): [list(c) for c in v] for k, v in sol.items()}})
        out["object_coverage"] = cov

    # ---- object movement (TAPVid-3D): format_training_samples on the tapvid_small fixture (OM-coord:317-404)
    with quiet:
        from spatial_engine.object_movement import single_object_movement_engine_coord as om
    tv = dict(np.load(os.path.join(HERE, "tapvid_small.npz")))
    T, N, _ = tv["tracks_xyz"].shape
    H, W = (int(v) for v in tv["image_hw"])
    eng3 = om.TwoFrameVideoQAEngine.__new__(om.TwoFrameVideoQAEngine)
    eng3.object_not_moving_threshold, eng3.camera_not_moving_threshold = 0.01, 0.01
    eng3.task_description, eng3.question_templates = ["t"], ["{x1} {y1}"]
    eng3.answer_templates = ["{total_distance} {x_value} {y_value} {z_value}"]
    eng3.question_type = "tapvid3d_total_distance"
    rng = np.random.default_rng(5)
    trip = [{"frame1": int(a), "frame2": int(b), "point_index": int(c)}
            for a, b, c in zip(rng.integers(0, T, 120), rng.integers(0, T, 120), rng.integers(0, N, 120)) if a != b]
    # float64 camera-frame tracks: the pinned numpy 1.26 promotes float32 + 1e-8 to float64 (see make_golden.py)
    with quiet:
        data = eng3.format_training_samples(trip, tv["fxfycxcy"], "video00003", tv["ref_world"].copy(),
                                            tv["tracks_xyz"].astype(np.float64), H, W, tv["w2c"])
    res = []
    for e in data:
        f1, f2, pt = (int(v) for v in e["id"].split("_")[1:4])
        nums = [int(v) for v in e["conversations"][1]["value"].replace("The point did not move. ", "").split()]
        res.append({"frame1": f1, "frame2": f2, "point": pt, "p1": [int(v) for v in e["p1"]], "p2": [int(v) for v in e["p2"]],
                    "total_distance": nums[0], "xyz": nums[1:], "gt_value": int(e["gt_value"]),
                    "point_moving": int(e["point_moving"]), "cam_moving": int(e["cam_moving"])})
    out["object_movement"] = res
    out["object_movement_requested"] = [[t["frame1"], t["frame2"], t["point_index"]] for t in trip]
    # rigid-body segmentation of the camera-frame tracks (OM-coord:49-105), as generate_qa_training_single_scene calls it
    groups = om.rigid_body_segmentation(tv["tracks_xyz"])
    out["rigid_body_groups"] = [[int(i) for i in grp] for grp in groups]
    out["rigid_body_groups_filtered"] = [[int(i) for i in grp] for grp in om.filter_large_groups(groups, min_size=5)]
    with open(os.path.join(HERE, "engines_small.json"), "w") as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
