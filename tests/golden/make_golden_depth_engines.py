"""Full-output goldens of the reference's four depth-perception engines, produced by the UNMODIFIED engine classes
(depth_estimation_{coor,dot}_engine.py, depth_comparison_{coor,dot}_engine.py) through their real constructors and
``generate_qa_training_data`` / ``generate_qa_eval_data`` on two synthetic disk scenes (build container only).

The engines' template lists are replaced by the short neutral lists below (several entries each, so that every
random.choice consumes the stream like a real draw); cv2.circle / cv2.putText are wrapped to log what the dot engines draw.
Stored: every sample dict, the position of Python's RNG stream afterwards, the draw log.

    python tests/golden/make_golden_depth_engines.py   ->  tests/golden/depth_engines.json.gz
"""
import contextlib
import gzip
import io
import json
import os
import pickle
import random
import sys
import tempfile

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from multi_spatialmllm_b200 import synth  # noqa: E402
from oracle import ref_layout, ref_shim  # noqa: E402

TEMPLATES = {
    "depth_estimation": {"task": ["<image>\nT0", "<image>\nT1", "<image>\nT2"], "questions": ["Q0 [ {x1} , {y1} ]", "Q1 {y1} {x1}"],
                         "answers": ["A0 {x1} {y1} `{depth}`", "A1 `{depth}`", "A2 {depth} {x1}"]},
    "depth_estimation_dot": {"task": ["<image>\nT0", "<image>\nT1"], "questions": ["Q0", "Q1", "Q2"],
                             "answers": ["A0 `{depth}`", "A1 {depth} mm"]},
    "depth_comparison": {"task": ["<image>\nT0", "<image>\nT1", "<image>\nT2"],
                         "closer_questions": ["CQ0 {x1} {y1} {x2} {y2}", "CQ1 {x2} {y2} {x1} {y1}"],
                         "farther_questions": ["FQ0 {x1} {y1} {x2} {y2}", "FQ1 {x2} {y2}", "FQ2 {y1}"],
                         "closer_answers": ["CA0 [ {correct_x} , {correct_y} ]", "CA1 {correct_y}"],
                         "farther_answers": ["FA0 [ {correct_x} , {correct_y} ]", "FA1 {correct_x}", "FA2"]},
    "depth_comparison_dot": {"task": ["<image>\nT0", "<image>\nT1"], "closer_questions": ["CQ0", "CQ1", "CQ2"],
                             "farther_questions": ["FQ0", "FQ1"], "closer_answers": ["CA0 `{correct_label}`", "CA1 {correct_label}"],
                             "farther_answers": ["FA0 `{correct_label}`", "FA1 {correct_label}", "FA2 {correct_label}!"]},
}

# (engine key, module, class, all_max_samples, max_n_points_per_image, seed, mode)
RUNS = [
    ("depth_estimation", "depth_estimation_coor_engine", "DepthEstimationCoorQAEngine", 11, 2, 4, "train"),
    ("depth_estimation", "depth_estimation_coor_engine", "DepthEstimationCoorQAEngine", 2, 1, 14, "eval"),
    ("depth_estimation_dot", "depth_estimation_dot_engine", "DepthEstimationDotQAEngine", 9, 1, 5, "train"),
    ("depth_comparison", "depth_comparison_coor_engine", "DepthComparisonCoorQAEngine", 30, 1, 6, "train"),
    ("depth_comparison", "depth_comparison_coor_engine", "DepthComparisonCoorQAEngine", 1, 1, 16, "eval"),
    ("depth_comparison_dot", "depth_comparison_dot_engine", "DepthComparisonDotQAEngine", 25, 1, 7, "train"),
]


def scenes_from_fixture():
    g = dict(np.load(os.path.join(HERE, "scannet_small.npz")))
    ids = [str(v) for v in g["image_ids"]]
    a = synth.SynthScene("scene0011_00", g["points"], g["instance_ids"], int(g["num_objects"]), g["axis_align"], g["c2w"],
                         g["K"], tuple(int(v) for v in g["image_hw"]), tuple(int(v) for v in g["depth_hw"]),
                         depth=g["depth"], image_ids=ids)
    keep = [0, 2, 3, 4, 6]                                      # a second, shorter scene (frame 4 has the -inf pose)
    b = synth.SynthScene("scene0012_00", g["points"], g["instance_ids"], int(g["num_objects"]), g["axis_align"], g["c2w"][keep],
                         g["K"], a.image_hw, a.depth_hw, depth=g["depth"][keep], image_ids=[ids[k] for k in keep])
    return [a, b]


def main():
    ref = ref_shim.import_reference()
    quiet = contextlib.redirect_stdout(io.StringIO())
    out = []
    with tempfile.TemporaryDirectory() as tmp, np.errstate(all="ignore"):
        root = os.path.join(tmp, "data", "scannet")
        scenes = scenes_from_fixture()
        info_path = ref_layout.write_split(scenes, root, name="scenes_val_info_i_D5.pkl")
        cwd = os.getcwd()
        os.chdir(tmp)                                            # the engines use the relative default roots data/scannet/...
        try:
            with quiet:
                h = ref.info_handler.SceneInfoHandler(info_path)
                vis = {s.scene_id: ref.mvi.process_scene(s.scene_id, h, os.path.join(tmp, "w.txt"))[1] for s in scenes}
            pkl = os.path.join(root, "scannet_instance_data", "val_visibility_info_D5.pkl")
            with open(pkl, "wb") as f:
                pickle.dump(vis, f)
            with quiet:
                ref.mvi.convert_pkl_to_parquet(pkl)
            for key, module, cls_name, n_max, npp, seed, mode in RUNS:
                with quiet:
                    mod = __import__(f"spatial_engine.depth_perception.{module}", fromlist=[cls_name])
                    eng = getattr(mod, cls_name)(scene_info_path=info_path, version_num="v1_0", all_max_samples=n_max,
                                                 image_output_dir=os.path.join(tmp, "out", key, "images"),
                                                 visibility_info_path=pkl.replace(".pkl", ".parquet"),
                                                 max_n_points_per_image=npp, warning_file=os.path.join(tmp, "warn.txt"))
                t = TEMPLATES[key]
                eng.task_description, eng.templates = t["task"], {k: v for k, v in t.items() if k != "task"}
                log, real_circle, real_text = [], cv2.circle, cv2.putText

                def circle(img, center, radius, color, thickness=1, *a, **k):
                    log.append(["circle", int(center[0]), int(center[1]), int(radius), [int(c) for c in color], int(thickness)])
                    return real_circle(img, center, radius, color, thickness, *a, **k)

                def text(img, s, org, font, scale, color, thickness=1, *a, **k):
                    log.append(["text", s, int(org[0]), int(org[1]), int(font), float(scale), [int(c) for c in color], int(thickness)])
                    return real_text(img, s, org, font, scale, color, thickness, *a, **k)

                mod.cv2.circle, mod.cv2.putText = circle, text
                random.seed(seed)
                try:
                    with quiet:
                        if mode == "train":
                            data = eng.generate_qa_training_data(os.path.join(tmp, "out", key), save_file=False)
                        else:
                            eng.generate_qa_eval_data(os.path.join(tmp, "out", key + "_eval"))
                            data = [json.loads(l) for l in open(os.path.join(tmp, "out", key + "_eval", f"{eng.task_name}.jsonl"))]
                finally:
                    mod.cv2.circle, mod.cv2.putText = real_circle, real_text
                after = random.random()
                img_dir = os.path.join(tmp, "out", key, "images")
                files = sorted(os.path.join(d, f)[len(img_dir) + 1:] for d, _, fs in os.walk(img_dir) for f in fs) if os.path.isdir(img_dir) else []
                out.append({"run": [key, cls_name, n_max, npp, seed, mode], "samples": json.loads(json.dumps(data)),
                            "random_after": after, "draw_log": log, "files": files})
                print(cls_name, mode, len(data), "samples,", len(log), "draw calls,", len(files), "files")
        finally:
            os.chdir(cwd)
    with gzip.open(os.path.join(HERE, "depth_engines.json.gz"), "wt") as f:
        json.dump({"templates": TEMPLATES, "runs": out}, f)


if __name__ == "__main__":
    main()
