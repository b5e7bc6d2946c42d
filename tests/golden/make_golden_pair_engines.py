"""Full-output goldens of the reference's pair-table engines and of its object-perception writer, produced by the
UNMODIFIED reference modules (build container only; needs /root/reference):

  camera_movement_engine_train_val.build_train_dataset / build_val_dataset                 (CME:269-350)
  visual_correspondence_qa_engine_coor_2_coor.build_train_dataset / build_val_dataset      (VC-coor:401-468)
  visual_correspondence_qa_engine_dot_2_multichoice.build_train_dataset                    (VC-dot:436-470)
  single_object_perception_engine.build_lwh_qa_samples                                     (SOPE:129-227)
  sample_dataframe                                                                          (CME:29-151) on random tables

on two synthetic disk scenes (``scenes()`` below: seeded synth scenes with oracle-rendered depth maps, rebuilt the same
way by the tests), with the reference's own L2 tables as inputs (make_visibility_info.process_scene,
calculate_frames_relations.process_scene + save_overlap_info).

Shims, all on the reference side and all for defects listed in SURVEY.md §2.3: ``get_image_size`` / ``image_height`` /
``image_width`` are set on the reference handler instance (the engines call them, the handler lacks them), and SOPE's
undefined ``TASK_DESCRIPTION`` is bound to a list.  Template lists are replaced by short neutral ones (several entries
each so that every random.choice is a real draw); cv2.circle / cv2.putText are wrapped to log what VC-dot draws.
``sample_dataframe`` is also run with ``groupby(observed=False)`` forced — the behaviour of the pandas 2.2.3 the
reference pins, where empty bins take part in the quota split — next to this container's pandas 3 default.

    python tests/golden/make_golden_pair_engines.py   ->  tests/golden/pair_engines.json.gz
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
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from multi_spatialmllm_b200 import synth  # noqa: E402
from oracle import geometry as og  # noqa: E402
from oracle import ref_layout, ref_shim  # noqa: E402

QTYPES = ("x_movement", "y_movement", "z_movement", "yaw_movement", "pitch_movement", "total_distance", "yaw_angle",
          "pitch_angle", "displacement_vector")
TEMPLATES = {
    "camera_movement": {"task": ["Image-1: <image>\nImage-2: <image>\nT0", "Image-1: <image>\nImage-2: <image>\nT1", "T2"],
                        "questions": {q: [f"Q0 {q}", f"Q1 {q}", f"Q2 {q}"] for q in QTYPES},
                        "answers": {q: ["A0 `{" + q + "}`", "A1 {" + q + "} {total_distance}"] for q in QTYPES}},
    "visual_correspondence": {"task": ["T0", "T1", "T2"], "questions": ["Q0 [ {x1} , {y1} ]", "Q1 {y1} {x1}"],
                              "answers": ["A0 [ {x2} , {y2} ]", "A1 {x2} {y2} {x1}", "A2 {y2}"]},
    "visual_correspondence_dot": {"task": ["T0", "T1"], "questions": ["Q0", "Q1", "Q2"],
                                  "answers": ["A0 `{correct_label}`", "A1 {correct_label}"]},
    "object_perception": {"task": ["T0", "T1", "T2"], "questions": ["Q0 {dimension} {object_category}", "Q1 {dimension}"],
                          "answers": ["A0 {dimension} `{value_mm}`", "A1 {object_category} {value_mm}", "A2 {value_mm}"]},
}
BINS = (6, 35, 1)


def scenes():
    """Two small scenes; the depth maps are the oracle's z-buffer of the scene's own points (+1 mm margin)."""
    out = []
    for seed, n_frames, sid in ((21, 30, "scene0021_00"), (22, 24, "scene0022_00")):
        s = synth.make_scene(seed, n_points=3000, n_frames=n_frames, n_objects=5, scene_id=sid,
                             invalid_pose_every=7 if seed == 21 else 0)
        pts = s.points.astype(np.float64)
        H, W = s.image_hw
        depth = np.zeros((n_frames, *s.depth_hw), np.uint16)
        with np.errstate(all="ignore"):
            for f in range(n_frames):
                if np.isfinite(s.c2w[f]).all():
                    E = og.aligned_pose(s.axis_align, s.c2w[f])
                    depth[f] = og.depth_to_u16(og.zbuffer_render(pts, s.K, E, H, W, *s.depth_hw), 1)
        s.depth = depth
        out.append(s)
    return out


def sampler_cases():
    """(seed, rows, all_overlap_samples, non_overlap_samples, bins): overlap tables with zeros, NaNs, values on bin edges,
    empty bins and bins both smaller and larger than their quota."""
    return [(0, 400, 60, 0, (6, 35, 1)), (1, 400, 200, 5, (6, 35, 1)), (2, 50, 30, 0, (6, 35, 1)), (3, 300, 45, 3, (0, 100, 10)),
            (4, 120, 1000, 100, (6, 35, 1)), (5, 30, 10, 0, (40, 50, 1)), (6, 500, 97, 0, (0, 100, 1))]


def sampler_table(seed, rows):
    rng = np.random.default_rng(1000 + seed)
    ov = rng.uniform(0, 60, rows) ** (1.0 if seed % 2 else 1.3) % 100
    ov[rng.random(rows) < 0.1] = 0.0
    ov[rng.random(rows) < 0.05] = np.nan
    edge = rng.random(rows) < 0.1
    ov[edge] = np.round(ov[edge])                                        # exactly on bin edges (6.0 and 35.0 included)
    ov[:3] = [6.0, 35.0, 35.000001]
    return pd.DataFrame({"scene_id": [f"s{k % 7}" for k in range(rows)], "image_id1": [f"{k:05d}" for k in range(rows)],
                         "image_id2": [f"{k + 1:05d}" for k in range(rows)], "overlap": ov, "distance": rng.random(rows),
                         "yaw": rng.uniform(-300, 300, rows), "pitch": rng.uniform(-80, 80, rows)})


def object_coverage_info(scene_list, seed=5):
    """A hand-made merged coverage file (SOPE:138-150 layout) over the scenes' real image ids: ints and str keys, an
    out-of-range k, a non-numeric key and an empty combination."""
    rng = random.Random(seed)
    info = {}
    for s in scene_list:
        ids = [i for f, i in enumerate(s.image_ids) if np.isfinite(s.c2w[f]).all()]
        info[s.scene_id] = {}
        for obj in range(s.num_objects):
            if obj == 3:
                continue
            d = {}
            for k in (1, 2, "3", 7, "x"):
                kk = k if isinstance(k, int) else (int(k) if k.isdigit() else 2)
                d[k] = [tuple(rng.sample(ids, kk)) for _ in range(rng.randint(0, 4))]
            d[2] = d[2] + [()]
            info[s.scene_id][obj] = d
    return info


def main():
    ref = ref_shim.import_reference()
    quiet = contextlib.redirect_stdout(io.StringIO())
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "spatial_engine", "camera_movement"))   # `from TEMPLATES import`
    with quiet:
        from spatial_engine.camera_movement import camera_movement_engine_train_val as cme
        from spatial_engine.visual_correspondence import visual_correspondence_qa_engine_coor_2_coor as vcc
        from spatial_engine.visual_correspondence import visual_correspondence_qa_engine_dot_2_multichoice as vcd
        from spatial_engine.object_perception import single_object_perception_engine as sope
    out = {"templates": TEMPLATES, "bins": list(BINS), "runs": [], "sampler": [], "object_perception": []}

    # ---- sample_dataframe alone (pure pandas / numpy RNG)
    real_groupby = pd.DataFrame.groupby
    for seed, rows, n_all, n_zero, bins in sampler_cases():
        df = sampler_table(seed, rows)
        df["row"] = np.arange(rows)
        for observed in (None, False):
            if observed is False:
                pd.DataFrame.groupby = lambda self, *a, **k: real_groupby(self, *a, **{**k, "observed": False})
            try:
                np.random.seed(seed)
                with quiet:
                    got = cme.sample_dataframe(df, n_all, n_zero, *bins)
                after = float(np.random.random())
            finally:
                pd.DataFrame.groupby = real_groupby
            out["sampler"].append({"case": [seed, rows, n_all, n_zero, list(bins)], "observed": observed is None,
                                   "rows": [int(v) for v in got["row"]], "np_random_after": after})
            print("sampler", seed, "observed" if observed is None else "all bins", len(got), "rows")

    with tempfile.TemporaryDirectory() as tmp, np.errstate(all="ignore"):
        root = os.path.join(tmp, "data", "scannet")
        scene_list = scenes()
        info_path = ref_layout.write_split(scene_list, root, name="scenes_train_val_info_i_D5.pkl")
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            with quiet:
                h = ref.info_handler.SceneInfoHandler(info_path)
                h.get_image_size = h.get_image_shape                               # reference defects, SURVEY.md §2.3
                h.image_height, h.image_width = h.get_image_shape(scene_list[0].scene_id)
                vis = {s.scene_id: ref.mvi.process_scene(s.scene_id, h, os.path.join(tmp, "w.txt"))[1] for s in scene_list}
                rel = {s.scene_id: ref.cfr.process_scene(s.scene_id, h, os.path.join(tmp, "w.txt"))[1] for s in scene_list}
            vis_pkl = os.path.join(root, "scannet_instance_data", "val_visibility_info_D5.pkl")
            with open(vis_pkl, "wb") as f:
                pickle.dump(vis, f)
            rel_parquet = os.path.join(tmp, "camera_info_D5.parquet")
            with quiet:
                ref.mvi.convert_pkl_to_parquet(vis_pkl)
                ref.cfr.save_overlap_info(rel, rel_parquet)
            table = pd.read_parquet(rel_parquet)
            in_bins = int(((table["overlap"] >= BINS[0]) & (table["overlap"] <= BINS[1])).sum())
            print(f"relations table: {len(table)} rows, {in_bins} inside the overlap bins")

            def run(name, mod, fn, seed, kwargs, log=None):
                random.seed(seed)
                np.random.seed(seed)
                od = os.path.join(tmp, "out", f"{name}_{seed}")
                os.makedirs(od, exist_ok=True)
                with quiet:
                    getattr(mod, fn)(parquet_path=rel_parquet, output_dir=od, scene_infos=h, overlap_min=BINS[0],
                                     overlap_max=BINS[1], interval=BINS[2], **kwargs)
                files = sorted(f for f in os.listdir(od) if f.endswith(".jsonl"))
                assert len(files) == 1, files
                data = [json.loads(l) for l in open(os.path.join(od, files[0]))]
                img_dir = os.path.join(od, "images")
                images = sorted(os.path.join(d, f)[len(img_dir) + 1:] for d, _, fs in os.walk(img_dir) for f in fs) if os.path.isdir(img_dir) else []
                warn = kwargs.get("warning_file")
                out["runs"].append({"engine": name, "fn": fn, "seed": seed,
                                    "kwargs": {k: v for k, v in kwargs.items() if k in ("qtype", "desired_count", "max_points_per_pair")},
                                    "file": files[0], "samples": data, "random_after": random.random(),
                                    "np_random_after": float(np.random.random()), "draw_log": log or [], "images": images,
                                    "warnings": open(warn).read() if warn and os.path.exists(warn) else ""})
                print(name, fn, seed, len(data), "samples,", len(log or []), "draw calls,", len(images), "images")

            # ---- camera movement
            t = TEMPLATES["camera_movement"]
            cme.TASK_DESCRIPTION, cme.QUESTION_TEMPLATES, cme.ANSWER_TEMPLATES = t["task"], t["questions"], t["answers"]
            run("camera_movement", cme, "build_train_dataset", 0, {"qtype": "displacement_vector", "desired_count": 40})
            run("camera_movement", cme, "build_val_dataset", 3, {"qtype": "yaw_angle", "desired_count": 25})
            run("camera_movement", cme, "build_train_dataset", 8, {"qtype": "x_movement", "desired_count": 1000})

            # ---- visual correspondence, coordinates (pickle visibility like the reference's main, VC-coor:471)
            t = TEMPLATES["visual_correspondence"]
            vcc.TASK_DESCRIPTION, vcc.TEMPLATES = t["task"], {k: v for k, v in t.items() if k != "task"}
            w = os.path.join(tmp, "vcc_warn.txt")
            run("visual_correspondence", vcc, "build_train_dataset", 1,
                {"desired_count": 45, "visibility_info_path": vis_pkl, "warning_file": w, "max_points_per_pair": 1})
            run("visual_correspondence", vcc, "build_train_dataset", 11,
                {"desired_count": 30, "visibility_info_path": vis_pkl, "warning_file": w, "max_points_per_pair": 3})
            run("visual_correspondence", vcc, "build_val_dataset", 12,
                {"desired_count": 1000, "visibility_info_path": vis_pkl, "warning_file": w, "max_points_per_pair": 1})

            # ---- visual correspondence, dots (parquet visibility like the reference's main, VC-dot:21)
            t = TEMPLATES["visual_correspondence_dot"]
            vcd.TASK_DESCRIPTION, vcd.TEMPLATES = t["task"], {k: v for k, v in t.items() if k != "task"}
            log, real_circle, real_text = [], cv2.circle, cv2.putText

            def circle(img, center, radius, color, thickness=1, *a, **k):
                log.append(["circle", int(center[0]), int(center[1]), int(radius), [int(c) for c in color], int(thickness)])
                return real_circle(img, center, radius, color, thickness, *a, **k)

            def text(img, s, org, font, scale, color, thickness=1, *a, **k):
                log.append(["text", s, int(org[0]), int(org[1]), int(font), float(scale), [int(c) for c in color], int(thickness)])
                return real_text(img, s, org, font, scale, color, thickness, *a, **k)

            cv2.circle, cv2.putText = circle, text
            try:
                run("visual_correspondence_dot", vcd, "build_train_dataset", 2,
                    {"desired_count": 35, "visibility_info_path": vis_pkl.replace(".pkl", ".parquet"),
                     "warning_file": os.path.join(tmp, "vcd_warn.txt"), "max_points_per_pair": 1}, log)
            finally:
                cv2.circle, cv2.putText = real_circle, real_text

            # ---- object perception writer
            t = TEMPLATES["object_perception"]
            sope.TASK_DESCRIPTION, sope.QUESTION_TEMPLATES, sope.ANSWER_TEMPLATES = t["task"], t["questions"], t["answers"]
            cov_pkl = os.path.join(tmp, "merged_train_object_coverage.pkl")
            with open(cov_pkl, "wb") as f:
                pickle.dump(object_coverage_info(scene_list), f)
            for dim, max_samples, seed in (("height", -1, 1), ("length", 4, 9), ("width", 2, 10)):
                od = os.path.join(tmp, "out", f"sope_{dim}")
                random.seed(seed)
                with quiet:
                    sope.build_lwh_qa_samples(h, cov_pkl, dim, "train", od, max_k=6, max_samples=max_samples)
                files = {f: [json.loads(l) for l in open(os.path.join(od, f))] for f in sorted(os.listdir(od))}
                out["object_perception"].append({"dimension": dim, "max_samples": max_samples, "seed": seed, "files": files,
                                                 "random_after": random.random()})
                print("object_perception", dim, {f: len(v) for f, v in files.items()})
        finally:
            os.chdir(cwd)
    with gzip.open(os.path.join(HERE, "pair_engines.json.gz"), "wt") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
