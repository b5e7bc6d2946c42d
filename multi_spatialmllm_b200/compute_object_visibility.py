"""Drop-in for spatial_engine/object_perception/compute_object_visibility.py — same function names and arguments.

  load_visibility_dict(parquet_file)                                          COV:60-70
  process_scene(scene_id, scene_info_handler, visibility_dict)                COV:72-152  -> (scene_id, result, warnings)
  process_split(split_name, scene_info_path, visibility_parquet_file, output_dir)   COV:154-194
  main()                                                                      COV:197-218

The per-(object, frame) ``len(set(json.loads(...)) & object_points_set)`` of the reference is the fused object-count
histogram of K1 (``msp_project_visibility_objects``): one launch per scene.  ``visibility_dict`` is accepted for signature
compatibility and not read — the masks are recomputed (they are bit-identical to the table, tests/test_gpu_parity.py).
"""
from __future__ import annotations

import pandas as pd

from . import object_perception, tables
from .tables import NONINFORMATIVE_DESC  # noqa: F401  (COV:58)


def load_visibility_dict(parquet_file):
    df = pd.read_parquet(parquet_file)
    return dict(zip(df["key"].tolist(), df["values"].tolist()))


def process_scene(scene_id, scene_info_handler, visibility_dict=None, device="cuda:0"):
    print(f"Processing scene {scene_id}.")
    warnings_list = []
    if scene_id not in scene_info_handler.infos:
        msg = f"[Warning] Scene {scene_id} not found in scene_info."
        print(msg)
        return scene_id, {"object_to_images": {}, "image_to_objects": {}}, [msg]
    sc, image_ids = tables.device_scene_from_handler(scene_info_handler, scene_id, device)
    result = tables.object_visibility_scene(scene_info_handler, scene_id, sc, image_ids, warnings_list)
    return scene_id, result, warnings_list


def process_split(split_name, scene_info_path, visibility_parquet_file, output_dir, device="cuda:0"):
    return object_perception.process_visibility_split(split_name, scene_info_path, visibility_parquet_file, output_dir, device)


def main(argv=None):
    return tables.main(["object_visibility"] + list(argv or []))


if __name__ == "__main__":
    import sys
    main(sys.argv[1:])
