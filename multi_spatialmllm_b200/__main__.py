"""``python -m multi_spatialmllm_b200 <reference script> [args...]`` — the reference's command lines by their own names.

Each name is the file name of a reference script (README.md:166-192 run them in this order); the arguments are the
drop-in's (``--help`` on any of them).  Without arguments the table below is printed.
"""
from __future__ import annotations

import sys

COMMANDS = {
    # reference script (without .py)                      -> (module, callable, leading argv, reference path)
    "batch_load_scannet_data": ("ingest", "load_main", [], "spatial_engine/utils/scannet_utils/batch_load_scannet_data.py"),
    "extract_posed_images": ("ingest", "extract_main", [], "spatial_engine/utils/scannet_utils/extract_posed_images.py"),
    "update_info_file_with_images": ("ingest", "update_info_main", [], "spatial_engine/utils/scannet_utils/update_info_file_with_images.py"),
    "make_visibility_info": ("tables", "main", ["visibility"], "spatial_engine/utils/scannet_utils/make_visibility_info.py"),
    "calculate_frames_relations": ("tables", "main", ["relations"], "spatial_engine/camera_movement/calculate_frames_relations.py"),
    "camera_movement_engine_train_val": ("pair_engines", "main", ["camera_movement"],
                                         "spatial_engine/camera_movement/camera_movement_engine_train_val.py"),
    "depth_estimation_coor_engine": ("engines", "main", ["depth_estimation_coor"], "spatial_engine/depth_perception/depth_estimation_coor_engine.py"),
    "depth_estimation_dot_engine": ("engines", "main", ["depth_estimation_dot"], "spatial_engine/depth_perception/depth_estimation_dot_engine.py"),
    "depth_comparison_coor_engine": ("engines", "main", ["depth_comparison_coor"], "spatial_engine/depth_perception/depth_comparison_coor_engine.py"),
    "depth_comparison_dot_engine": ("engines", "main", ["depth_comparison_dot"], "spatial_engine/depth_perception/depth_comparison_dot_engine.py"),
    "visual_correspondence_qa_engine_coor_2_coor": ("pair_engines", "main", ["visual_correspondence_coor_2_coor"],
                                                    "spatial_engine/visual_correspondence/visual_correspondence_qa_engine_coor_2_coor.py"),
    "visual_correspondence_qa_engine_dot_2_multichoice": ("pair_engines", "main", ["visual_correspondence_dot_2_multichoice"],
                                                          "spatial_engine/visual_correspondence/visual_correspondence_qa_engine_dot_2_multichoice.py"),
    "compute_object_visibility": ("tables", "main", ["object_visibility"], "spatial_engine/object_perception/compute_object_visibility.py"),
    "single_object_coverage_finder": ("object_perception", "coverage_main", [], "spatial_engine/object_perception/single_object_coverage_finder.py"),
    "merge_object_coverage": ("object_perception", "merge_main", [], "spatial_engine/object_perception/merge_object_coverage.py"),
    "single_object_perception_engine": ("object_perception", "main", [], "spatial_engine/object_perception/single_object_perception_engine.py"),
    "single_object_movement_engine_coord": ("tapvid", "main", ["coord"], "spatial_engine/object_movement/single_object_movement_engine_coord.py"),
    "single_object_movement_engine_dot": ("tapvid", "main", ["dot"], "spatial_engine/object_movement/single_object_movement_engine_dot.py"),
    "regen": ("regen", "main", [], "(all of the above for a sharded run, batch writers)"),
}


def usage() -> str:
    lines = [__doc__.strip(), "", f"{'command':52s} replaces"]
    lines += [f"{name:52s} {ref}" for name, (_, _, _, ref) in COMMANDS.items()]
    return "\n".join(lines)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help", "help"):
        print(usage())
        return 0
    name = argv[0][:-3] if argv[0].endswith(".py") else argv[0]
    if name not in COMMANDS:
        print(f"unknown command {argv[0]!r}\n\n{usage()}", file=sys.stderr)
        return 2
    module, fn, lead, _ = COMMANDS[name]
    import importlib
    mod = importlib.import_module(f"{__package__}.{module}")
    result = getattr(mod, fn)(lead + argv[1:])
    return result if isinstance(result, int) else 0


if __name__ == "__main__":
    raise SystemExit(main())
