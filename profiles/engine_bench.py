"""Rows per second of the L3 QA engines on a synthetic disk split in the reference's layout (config-1-shaped scenes:
81 369 points, 1296x968 JPEG + 640x480 PNG depth per frame).

    python profiles/engine_bench.py ours        (GPU box)      -> this repo's drop-in engines, whole split
    python profiles/engine_bench.py reference   (build container, needs /root/reference) -> the unmodified reference
                                                 functions on a bounded sample of the same rows, one process

Both sides build the same split (seeded synth scenes, depth maps from the oracle's z-buffer so that no GPU is needed to
write it) and read the same L2 tables' definitions; what is timed is the engines' sample-building loop
(build_training_sample per row / generate_qa_training_single_scene), not the table precompute.
"""
import contextlib
import io
import json
import os
import pickle
import random
import sys
import tempfile
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_spatialmllm_b200 import synth  # noqa: E402
from oracle import geometry as og  # noqa: E402
from oracle import ref_layout  # noqa: E402

N_SCENES, N_POINTS, N_FRAMES = int(os.environ.get("EB_SCENES", 4)), 81_369, int(os.environ.get("EB_FRAMES", 48))
BINS = (6, 35, 1)


def build_split(tmp):
    scenes = []
    for k in range(N_SCENES):
        s = synth.make_scene(100 + k, n_points=N_POINTS, n_frames=N_FRAMES, n_objects=12, scene_id=f"scene{100 + k:04d}_00")
        pts = s.points.astype(np.float64)
        H, W = s.image_hw
        s.depth = np.stack([og.depth_to_u16(og.zbuffer_render(pts, s.K, og.aligned_pose(s.axis_align, s.c2w[f]), H, W, *s.depth_hw), 1)
                            for f in range(N_FRAMES)])
        scenes.append(s)
    root = os.path.join(tmp, "data", "scannet")
    return scenes, root, ref_layout.write_split(scenes, root, name="scenes_train_val_info_i_D5.pkl")


def ours(tmp):
    import torch
    from multi_spatialmllm_b200 import engines, pair_engines, tables
    from multi_spatialmllm_b200.handler.info_handler import SceneInfoHandler
    scenes, root, info = build_split(tmp)
    h = SceneInfoHandler(info, posed_images_root=os.path.join(root, "posed_images"),
                         instance_data_root=os.path.join(root, "scannet_instance_data"))
    out = {"side": "ours", "gpu": torch.cuda.get_device_name(0), "scenes": N_SCENES, "points": N_POINTS, "frames": N_FRAMES}
    quiet = contextlib.redirect_stdout(io.StringIO())
    t = time.time()
    with quiet:
        tables.run_relations_split(h, os.path.join(tmp, "cam.parquet"))
    out["relations_table_s"] = round(time.time() - t, 3)
    df = pd.read_parquet(os.path.join(tmp, "cam.parquet"))
    out["table_rows"] = len(df)
    out["rows_in_bins"] = int(((df["overlap"] >= BINS[0]) & (df["overlap"] <= BINS[1])).sum())
    cache = pair_engines.SceneCache(h)
    from multi_spatialmllm_b200 import annotate                               # warm the nvJPEG worker pool (a handle per thread)
    blank = open(h.get_image_path(scenes[0].scene_id, scenes[0].image_ids[0]), "rb").read()
    annotate.write_annotated([(k, f"warm/{k}.jpg", []) for k in range(8)], lambda k: blank, os.path.join(tmp, "warm"))
    for name, cls, kw in (("visual_correspondence_coor", pair_engines.VisualCorrespondenceCoorEngine, {"max_points_per_pair": 1}),
                          ("visual_correspondence_dot", pair_engines.VisualCorrespondenceDotEngine, {"max_points_per_pair": 1}),
                          ("camera_movement", pair_engines.CameraMovementEngine, {"qtype": "displacement_vector"})):
        runs = []
        for rep in range(2):                                 # run 0 also pays the scene uploads of the shared cache
            random.seed(1)
            np.random.seed(1)
            eng = cls(h, cache=cache)
            od = os.path.join(tmp, f"{name}_{rep}")
            os.makedirs(od, exist_ok=True)
            torch.cuda.synchronize()
            t = time.time()
            with quiet:
                if name == "camera_movement":
                    s = eng.build_train_dataset(os.path.join(tmp, "cam.parquet"), od, kw["qtype"], 10 ** 7, *BINS)
                else:
                    s = eng.build_train_dataset(os.path.join(tmp, "cam.parquet"), od, 10 ** 7, *BINS, warning_file=os.path.join(od, "w.txt"), **kw)
            runs.append(time.time() - t)
        out[name] = {"samples": len(s), "seconds_first": round(runs[0], 3), "seconds": round(runs[1], 3),
                     "rows_per_s": round(len(s) / runs[1], 1)}
    for name, cls in (("depth_estimation_coor", engines.DepthEstimationCoorQAEngine), ("depth_comparison_coor", engines.DepthComparisonCoorQAEngine),
                      ("depth_estimation_dot", engines.DepthEstimationDotQAEngine)):
        runs = []
        for rep in range(2):
            random.seed(4)
            eng = cls(h, "v1_0", N_SCENES * N_FRAMES, os.path.join(tmp, f"{name}_{rep}", "images"), None, 1, os.path.join(tmp, "w.txt"))
            t = time.time()
            with quiet:
                s = eng.generate_qa_training_data(os.path.join(tmp, name), save_file=False)
            runs.append(time.time() - t)
        out[name] = {"samples": len(s), "seconds_first": round(runs[0], 3), "seconds": round(runs[1], 3),
                     "rows_per_s": round(len(s) / runs[1], 1),
                     "note": "every run uploads each scene again (one SceneSession per scene) and builds its per-frame tables"}
    return out


def reference(tmp, sample_rows=60):
    from oracle import ref_shim
    ref = ref_shim.import_reference()
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "spatial_engine", "camera_movement"))
    quiet = contextlib.redirect_stdout(io.StringIO())
    with quiet:
        from spatial_engine.camera_movement import camera_movement_engine_train_val as cme
        from spatial_engine.visual_correspondence import visual_correspondence_qa_engine_coor_2_coor as vcc
        from spatial_engine.visual_correspondence import visual_correspondence_qa_engine_dot_2_multichoice as vcd
        from spatial_engine.depth_perception import depth_estimation_coor_engine as de
    scenes, root, info = build_split(tmp)
    out = {"side": "reference", "cores_used": 1, "host_cores": os.cpu_count(), "scenes": N_SCENES, "points": N_POINTS, "frames": N_FRAMES}
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        with quiet, np.errstate(all="ignore"):
            h = ref.info_handler.SceneInfoHandler(info)
            h.get_image_size = h.get_image_shape
            h.image_height, h.image_width = h.get_image_shape(scenes[0].scene_id)
            one = scenes[0].scene_id
            t = time.time()
            vis = {one: ref.mvi.process_scene(one, h, os.path.join(tmp, "w.txt"))[1]}
            out["visibility_table_s_per_scene"] = round(time.time() - t, 2)
            t = time.time()
            rel = {one: ref.cfr.process_scene(one, h, os.path.join(tmp, "w.txt"))[1]}
            out["relations_table_s_per_scene"] = round(time.time() - t, 2)
            ref.cfr.save_overlap_info(rel, os.path.join(tmp, "cam.parquet"))
            df = pd.read_parquet(os.path.join(tmp, "cam.parquet"))
            df = df[(df["overlap"] >= BINS[0]) & (df["overlap"] <= BINS[1])].head(sample_rows).reset_index(drop=True)
            vis_parquet = {f"{one}:image_to_points:{k}": json.dumps(v) for k, v in vis[one]["image_to_points"].items()}
            for name, fn in (("visual_correspondence_coor", lambda r, i: vcc.build_training_sample(h, r, i, vis, os.path.join(tmp, "w.txt"))),
                             ("visual_correspondence_dot", lambda r, i: vcd.build_training_sample(h, r, i, vis_parquet, os.path.join(tmp, "w.txt"), image_output_dir=os.path.join(tmp, "img"))),
                             ("camera_movement", lambda r, i: cme.build_training_sample(h, r, i, "displacement_vector"))):
                random.seed(1)
                t = time.time()
                n = sum(fn(df.iloc[i], i) is not None for i in range(len(df)))
                dt = time.time() - t
                out[name] = {"samples": n, "seconds": round(dt, 3), "rows_per_s": round(len(df) / dt, 1)}
            pkl = os.path.join(root, "scannet_instance_data", "val_visibility_info_D5.pkl")
            with open(pkl, "wb") as f:
                pickle.dump(vis, f)
            ref.mvi.convert_pkl_to_parquet(pkl)
            eng = de.DepthEstimationCoorQAEngine(scene_info_path=info, version_num="v1_0", all_max_samples=-1, image_output_dir=None,
                                                 visibility_info_path=pkl.replace(".pkl", ".parquet"), max_n_points_per_image=1,
                                                 warning_file=os.path.join(tmp, "w.txt"))
            eng.max_samples = N_FRAMES
            random.seed(4)
            t = time.time()
            s = eng.generate_qa_training_single_scene(one)
            dt = time.time() - t
            out["depth_estimation_coor"] = {"samples": len(s), "seconds": round(dt, 3), "rows_per_s": round(len(s) / dt, 1)}
    finally:
        os.chdir(cwd)
    return out


if __name__ == "__main__":
    side = sys.argv[1] if len(sys.argv) > 1 else "ours"
    with tempfile.TemporaryDirectory() as tmp:
        res = ours(tmp) if side == "ours" else reference(tmp)
    print(json.dumps(res))
