"""Drop-in for spatial_engine/utils/scannet_utils/make_visibility_info.py — same function names and arguments.

  process_scene(scene_id, scene_infos, warning_file)   MVI:75-125   -> (scene_id, {"image_to_points", "point_to_images"})
  convert_pkl_to_parquet(pkl_file)                      MVI:38-73
  run_split(scene_info_path, output_file, warning_file, num_workers=8)   MVI:128-177
  main()                                                MVI:179-214

The frame loop of ``process_scene`` (project every point, bounds, depth test: IH:313-386 per frame) is one K1 launch over
all frames of the scene; the lists are read off the bit masks.  ``run_split`` writes the parquet the reference's readers
expect (keys with ``:`` and JSON text, MVI:59-65 — the keys the reference's own run_split writes, MVI:163-172, use ``,`` and
cannot be read back by IH:510-513), one row group per scene, formatted on the device.
"""
from __future__ import annotations

import os
import pickle


from . import engine, tables
from .handler.info_handler import SceneInfoHandler


def process_scene(scene_id, scene_infos, warning_file, device="cuda:0"):
    print(f"[process_scene] Start: {scene_id}")
    sc, image_ids = tables.device_scene_from_handler(scene_infos, scene_id, device)
    result = engine.process_scene_visibility(sc, image_ids)
    for image_id, points in result["image_to_points"].items():                   # MVI:110-113
        if len(points) == 0:
            with open(warning_file, "a") as f:
                f.write(f"[Warning] {scene_id}: {image_id} has no in-bound points.\n")
    print(f"[process_scene] Done: {scene_id}")
    return scene_id, result


def convert_pkl_to_parquet(pkl_file):
    with open(pkl_file, "rb") as f:
        scene_visibility_dict = pickle.load(f)
    parquet_file = pkl_file.replace(".pkl", ".parquet")
    n = tables.save_visibility_parquet(scene_visibility_dict, parquet_file)
    print(f"Converted {pkl_file} to {parquet_file}. The file has {n} items in total.")


def run_split(scene_info_path, output_file, warning_file, num_workers=8, device="cuda:0"):
    """``num_workers`` is accepted for signature compatibility: one process drives one GPU (use ``torchrun`` +
    ``python -m multi_spatialmllm_b200.tables visibility`` for several)."""
    scene_infos = scene_info_path if isinstance(scene_info_path, SceneInfoHandler) else SceneInfoHandler(scene_info_path)
    os.makedirs(os.path.dirname(output_file) or ".", exist_ok=True)
    res = tables.run_visibility_split(scene_infos, output_file, device=device, warning_file=warning_file)
    print(f"[run_split] Done. Wrote {res['rows']} entries to {output_file}")
    return res


def main(argv=None):
    return tables.main(["visibility"] + list(argv or []))


if __name__ == "__main__":
    import sys
    main(sys.argv[1:])
