"""Golden for the TAPVid-3D per-video sample generation, produced by the UNMODIFIED reference
(single_object_movement_engine_coord.TwoFrameVideoQAEngine.generate_qa_training_single_scene, OM-coord:406-575) on seeded
synthetic clips written as TAPVid-3D npz files (build container only).

The clips store their camera-frame tracks as float64 holding float32-representable values: under the reference's pinned
numpy 1.26.4 ``float32 + 1e-8`` promotes to float64 (OM-coord:303-304), and float64 inputs pin exactly that semantics under
the numpy 2.x of this container (see make_golden.py).  The engine's three template lists are replaced by the short neutral
lists below (several entries each, so that every random.choice consumes the stream like a real draw): the reference's own
phrasings are CC BY-NC and stay out of this repository, while draw order, formatting and the RNG position stay pinned.

    python tests/golden/make_golden_tapvid_sampling.py   ->  tests/golden/tapvid_sampling.json.gz
"""
import contextlib
import gzip
import json
import os
import random
import sys
import tempfile

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from multi_spatialmllm_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

TEMPLATES = {"task": ["Image-1: <image>\nImage-2: <image>\nT0", "Image-1: <image>\nImage-2: <image>\nT1", "T2 <image> <image>"],
             "questions": ["Q0 [ {x1} , {y1} ]", "Q1 {y1} / {x1}"],
             "answers": ["A0 `{total_distance}` [ {x_value} , {y_value} , {z_value} ]", "A1 {z_value} {y_value} {x_value} {total_distance}",
                         "A2 {total_distance}", "A3 {x_value}", "A4 {y_value} {z_value}"]}

CASES = [  # (video seed, frames, tracks, groups, question type, npoints_per_group, augment_ratio, random seed)
    (11, 30, 60, 3, "tapvid3d_total_distance", 5, 1.0, 0),
    (12, 24, 48, 2, "tapvid3d_displacement_vector", 3, 0.5, 7),
]


def write_npz(path, v, image_hw=(36, 64)):
    ok, jpg = cv2.imencode(".jpg", np.zeros((*image_hw, 3), np.uint8))
    frames = np.array([jpg.tobytes()] * v.tracks_xyz.shape[0], dtype=object)
    np.savez(path, images_jpeg_bytes=frames, fx_fy_cx_cy=v.fxfycxcy, tracks_XYZ=v.tracks_xyz.astype(np.float64),
             visibility=v.visibility, queries_xyt=np.zeros((v.tracks_xyz.shape[1], 3)), extrinsics_w2c=v.w2c)


def make_clip(seed, T, N, G):
    v = synth.make_video(seed, n_frames=T, n_tracks=N, n_groups=G, image_hw=(36, 64))
    v.tracks_xyz[:, : N // 6] = v.tracks_xyz[:1, : N // 6]     # a few tracks frozen in camera coordinates
    return v


def main():
    ref_shim.import_reference()
    quiet = contextlib.redirect_stdout(open(os.devnull, "w"))
    with quiet:
        from spatial_engine.object_movement import single_object_movement_engine_coord as om
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        for (vs, T, N, G, qt, npg, ratio, rs) in CASES:
            v = make_clip(vs, T, N, G)
            path = os.path.join(tmp, f"{v.video_id}.npz")
            write_npz(path, v)
            eng = om.TwoFrameVideoQAEngine(qt, "adt")
            eng.task_description, eng.question_templates, eng.answer_templates = (TEMPLATES["task"], TEMPLATES["questions"],
                                                                                 TEMPLATES["answers"])
            random.seed(rs)
            with quiet:
                data = eng.generate_qa_training_single_scene(path, npoints_per_group=npg, npairs_per_bin=1e8,
                                                             img_output_dir=os.path.join(tmp, "img"), augment=True,
                                                             augment_ratio=ratio)
            after = random.random()                             # the RNG position after the call is part of the contract
            rows = [{"id": e["id"], "gt_value": e["gt_value"], "p1": list(e["p1"]), "p2": list(e["p2"]),
                     "point_moving": e["point_moving"], "cam_moving": e["cam_moving"], "image": e["image"],
                     "hw": [e["height_list"][0], e["width_list"][0]],
                     "conversations": e["conversations"]} for e in data]
            out.append({"case": [vs, T, N, G, qt, npg, ratio, rs], "samples": rows, "random_after": after})
            print(qt, len(rows), "samples;", sum(r["point_moving"] == 0 for r in rows), "static")
    # ---- dot variant (single_object_movement_engine_dot.py): same sampling, first image annotated; cv2.circle is wrapped
    #      (the library call, not the reference) to record centre / radius / colour of every annotation in order
    with quiet:
        from spatial_engine.object_movement import single_object_movement_engine_dot as omd
    dot_cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for (vs, T, N, G, qt, npg, ratio, rs) in [(13, 20, 40, 2, "tapvid3d_displacement_vector", 2, 0.3, 3)]:
            v = make_clip(vs, T, N, G)
            path = os.path.join(tmp, f"{v.video_id}.npz")
            write_npz(path, v)
            eng = omd.TwoFrameVideoQAEngineDot(qt, "adt")
            eng.task_description, eng.question_templates, eng.answer_templates = (TEMPLATES["task"], TEMPLATES["questions"],
                                                                                 TEMPLATES["answers"])
            circles, real_circle = [], cv2.circle

            def spy(img, center, radius, color, thickness=1, *a, **k):
                circles.append([int(center[0]), int(center[1]), int(radius), [int(c) for c in color], int(thickness)])
                return real_circle(img, center, radius, color, thickness, *a, **k)

            omd.cv2.circle = spy
            random.seed(rs)
            try:
                with quiet:
                    data = eng.generate_qa_training_single_scene(path, os.path.join(tmp, "base"), npoints_per_group=npg,
                                                                 npairs_per_bin=1e8, img_output_dir=os.path.join(tmp, "ann"),
                                                                 augment=True, augment_ratio=ratio)
            finally:
                omd.cv2.circle = real_circle
            after = random.random()
            rows = [{"id": e["id"], "gt_value": e["gt_value"], "p1": list(e["p1"]), "p2": list(e["p2"]),
                     "point_moving": e["point_moving"], "cam_moving": e["cam_moving"], "image": e["image"],
                     "hw": [e["height_list"][0], e["width_list"][0]], "conversations": e["conversations"]} for e in data]
            files = sorted(os.listdir(os.path.join(tmp, "ann", v.video_id)))
            dot_cases.append({"case": [vs, T, N, G, qt, npg, ratio, rs], "samples": rows, "random_after": after,
                              "circles": circles, "files": files})
            print("dot", qt, len(rows), "samples;", len(circles), "annotated images;", len(files), "files")
    with gzip.open(os.path.join(HERE, "tapvid_sampling.json.gz"), "wt") as f:
        json.dump({"templates": TEMPLATES, "cases": out, "dot_cases": dot_cases}, f)


if __name__ == "__main__":
    main()
