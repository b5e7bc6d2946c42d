"""Drop-in object_perception chain: visibility -> coverage search -> merge -> QA samples.

  compute_object_visibility.process_split / main          COV:154-218    -> process_visibility_split
  single_object_coverage_finder.process_split_objects     SOCF:287-350   -> process_split_objects
  single_object_coverage_finder.main (--split/--start/--end shards, find_object_coverage.sh) SOCF:353-422 -> coverage_main
  merge_object_coverage.merge_dimension / merge_split      merge_object_coverage.py:17-83 -> merge_dimension / merge_split
  single_object_perception_engine.build_lwh_qa_samples     SOPE:129-227   -> build_lwh_qa_samples
  single_object_perception_engine.build_train_and_val_datasets  SOPE:229-268 -> build_train_and_val_datasets

File names, pickle layouts and the consumption of Python's ``random`` stream are the reference's.  The geometry — per
(frame, object) visible counts and per-frame extents of each object's visible points — is one K1 launch and one
``msp_object_extents`` launch per scene (tables.object_visibility_scene / tables.object_coverage_scene); the reference
decodes a JSON list and builds a Python set per (object, frame) and a union mask per search node.

The published SOPE reads its task lines from a name it never defines (``TASK_DESCRIPTION``, SOPE:193; the list is bound
to ``ASK_DESCRIPTION``, SOPE:25) and stops with a NameError on the first sample; ``load_reference_templates`` accepts
either spelling.  Phrasings are not vendored (CC BY-NC).
"""
from __future__ import annotations

import glob
import json
import os
import pickle
import random
import re
import shutil

from . import streams, tables
from .handler.info_handler import SceneInfoHandler

NEUTRAL_TEMPLATES = {"task": ["Use the images to estimate the size of the object."],
                     "questions": ["What is the {dimension} of the {object_category} in millimetres?"],
                     "answers": ["`{value_mm}`"]}


def load_reference_templates(reference_root: str) -> dict:
    path = os.path.join(reference_root, "spatial_engine/object_perception/single_object_perception_engine.py")
    lit = streams.Templates._literals
    try:
        task = lit(path, ("TASK_DESCRIPTION",))["TASK_DESCRIPTION"]
    except KeyError:
        task = lit(path, ("ASK_DESCRIPTION",))["ASK_DESCRIPTION"]
    t = lit(path, ("QUESTION_TEMPLATES", "ANSWER_TEMPLATES"))
    return {"task": task, "questions": t["QUESTION_TEMPLATES"], "answers": t["ANSWER_TEMPLATES"]}


def _handler(scene_info):
    return scene_info if isinstance(scene_info, SceneInfoHandler) else SceneInfoHandler(scene_info)


# ------------------------------------------------------------------------------------------------
# L2 tables
# ------------------------------------------------------------------------------------------------
def process_visibility_split(split_name, scene_info_path, visibility_parquet_file, output_dir, device="cuda:0"):
    """COV:154-194.  ``visibility_parquet_file`` is accepted for signature compatibility (the masks are recomputed)."""
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "warning.txt"), "w") as wf:
        wf.write("")
    out = tables.run_object_visibility_split(_handler(scene_info_path), os.path.join(output_dir, "object_visibility.pkl"),
                                             device=device, warning_file=os.path.join(output_dir, "warning.txt"))
    print(f"Finished processing split '{split_name}'.")
    return out


def _coverage_of_scenes(scene_info_handler, scene_ids, object_visibility_dict, device):
    res = {"height": {}, "length": {}, "width": {}}
    for scene_id in scene_ids:
        print(f"Processing scene {scene_id} for object coverage.")
        sc, image_ids = tables.device_scene_from_handler(scene_info_handler, scene_id, device)
        scene_result = tables.object_coverage_scene(scene_info_handler, scene_id, sc, image_ids, object_visibility_dict[scene_id])
        if scene_result:                                                          # SOCF:323-330
            for dim in res:
                res[dim][scene_id] = {obj: r[dim] for obj, r in scene_result.items()}
    return res


def _dump(obj, path):
    with open(path, "wb") as f:
        pickle.dump(obj, f)


def _load(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def process_split_objects(split_name, scene_info_path, visibility_parquet_file, object_visibility_file, output_dir,
                          device="cuda:0"):
    """SOCF:287-350: three files ``{split}_object_coverage_{height,length,width}.pkl`` for the whole split."""
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "coverage_finding_warning_objects.txt"), "w") as wf:
        wf.write("")
    h = _handler(scene_info_path)
    res = _coverage_of_scenes(h, h.get_all_scene_ids(), _load(object_visibility_file), device)
    for dim, r in res.items():
        _dump(r, os.path.join(output_dir, f"{split_name}_object_coverage_{dim}.pkl"))
    print(f"Finished processing split '{split_name}' for object coverage.")
    return res


COVERAGE_SPLITS = {                                                              # SOCF:364-377
    "val": {"scene_info_path": "data/scannet/scannet_instance_data/scenes_val_info_i_D5.pkl",
            "visibility_parquet_file": "data/scannet/scannet_instance_data/val_visibility_info_D5.parquet",
            "object_visibility_file": "evaluation_data/object_perception/object_visibility.pkl",
            "output_dir": "evaluation_data/object_perception"},
    "train": {"scene_info_path": "data/scannet/scannet_instance_data/scenes_train_info_i_D5.pkl",
              "visibility_parquet_file": "data/scannet/scannet_instance_data/train_visibility_info_D5.parquet",
              "object_visibility_file": "training_data/object_perception/object_visibility.pkl",
              "output_dir": "training_data/object_perception"},
}


def coverage_main(argv=None, splits=None, device="cuda:0"):
    """SOCF:353-422: one shard ``[start, end)`` of a split's scenes -> ``<output_dir>/<split>_<start>_<end>/
    <split>_object_coverage_<dim>_<start>_<end>.pkl``.  Seeds ``random`` with 0 like the reference module (SOCF:35)."""
    import argparse
    ap = argparse.ArgumentParser(description="Process object coverage for a given split and scene index range.")
    ap.add_argument("--split", type=str, required=True)
    ap.add_argument("--start", type=int, required=True)
    ap.add_argument("--end", type=int, default=None)
    a = ap.parse_args(argv)
    splits = splits or COVERAGE_SPLITS
    if a.split not in splits:
        raise ValueError("Invalid split. Choose train or val.")
    cfg = dict(splits[a.split])
    out_dir = os.path.join(cfg["output_dir"], f"{a.split}_{a.start}_{a.end}")
    os.makedirs(out_dir, exist_ok=True)
    random.seed(0)
    h = _handler(cfg["scene_info_path"])
    selected = h.get_all_scene_ids()[a.start:a.end]
    print(f"Processing scenes from index {a.start} to {a.end} (total {len(selected)}) in split {a.split}.")
    res = _coverage_of_scenes(h, selected, _load(cfg["object_visibility_file"]), device)
    for dim, r in res.items():
        _dump(r, os.path.join(out_dir, f"{a.split}_object_coverage_{dim}_{a.start}_{a.end}.pkl"))
    return res


def merge_dimension(split, base_dir, dimension):
    """merge_object_coverage.py:17-64: union of the shard files of one dimension, shards in ascending start order."""
    merged = {}
    pattern = re.compile(fr"{split}_(\d+)_(\d+)")
    shards = []
    for d in os.listdir(base_dir):
        m = pattern.match(d)
        if os.path.isdir(os.path.join(base_dir, d)) and d.startswith(f"{split}_") and m:
            shards.append((int(m.group(1)), d))
    for _, d in sorted(shards, key=lambda s: s[0]):
        for file in glob.glob(os.path.join(base_dir, d, f"{split}_object_coverage_{dimension}_*_*.pkl")):
            merged.update(_load(file))
    return merged


def merge_split(split, base_dir, output_dir):
    """merge_object_coverage.py:66-83 -> ``merged_{split}_object_coverage_{dim}.pkl``."""
    out = {}
    for dim in ("height", "length", "width"):
        out[dim] = merge_dimension(split, base_dir, dim)
        _dump(out[dim], os.path.join(output_dir, f"merged_{split}_object_coverage_{dim}.pkl"))
    return out


def merge_main(argv=None):
    """merge_object_coverage.py:85-97: merge the shard files of both splits under the reference's default directories."""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--train_base", default="training_data/object_perception")
    ap.add_argument("--val_base", default="evaluation_data/object_perception")
    a = ap.parse_args(argv)
    return {"train": merge_split("train", a.train_base, a.train_base), "val": merge_split("val", a.val_base, a.val_base)}


# ------------------------------------------------------------------------------------------------
# L3 samples
# ------------------------------------------------------------------------------------------------
def build_lwh_qa_samples(scene_info_handler, dimension_info_path, dimension_name, split, output_dir, max_k=6, max_samples=-1,
                         templates=None, rnd=random):
    """SOPE:129-227: one sample per image combination, written per combination size k to
    ``object_perception_{dimension}_k{k}_{split}_{max_samples}.jsonl``; returns {k: [samples]}."""
    tp = templates or NEUTRAL_TEMPLATES
    print(f"Processing dimension: {dimension_name}, split: {split}")
    dim_info = _load(dimension_info_path)
    os.makedirs(output_dir, exist_ok=True)
    getter = {"height": scene_info_handler.get_object_height, "length": scene_info_handler.get_object_length,
              "width": scene_info_handler.get_object_width}.get(dimension_name)
    H, W = scene_info_handler.image_height, scene_info_handler.image_width
    samples_by_k = {k: [] for k in range(1, max_k + 1)}
    for scene_id, obj_dict in dim_info.items():
        for object_id, k_dict in obj_dict.items():
            val_mm = int(round((getter(scene_id, object_id) if getter else 0.0) * 1000))
            category = scene_info_handler.get_object_raw_category(scene_id, object_id)
            for k_str, combos in k_dict.items():
                try:
                    k_val = int(k_str)
                except (TypeError, ValueError):
                    continue
                if k_val < 1 or k_val > max_k:
                    continue
                for combo_idx, combo in enumerate(combos):
                    if not combo:
                        continue
                    combo = list(combo)
                    rnd.shuffle(combo)
                    prefix = "\n".join(f"Image-{i}: <image>" for i in range(1, len(combo) + 1))
                    task_line = rnd.choice(tp["task"])
                    question = rnd.choice(tp["questions"]).format(dimension=dimension_name, object_category=category)
                    answer = rnd.choice(tp["answers"]).format(dimension=dimension_name, value_mm=val_mm, object_category=category)
                    samples_by_k[k_val].append({
                        "id": f"{scene_id}_{object_id}_{k_val}_{combo_idx}", "image": [f"{scene_id}/{img}.jpg" for img in combo],
                        "conversations": [{"from": "human", "value": f"{prefix}\n{task_line}\n{question}"},
                                          {"from": "gpt", "value": answer}],
                        "height_list": [H] * len(combo), "width_list": [W] * len(combo),
                        "question_type": f"object_perception_{dimension_name}_estimation", "gt_value": val_mm})
    for k in range(1, max_k + 1):
        if not samples_by_k[k]:
            continue
        if max_samples > 0 and len(samples_by_k[k]) > max_samples:
            samples_by_k[k] = rnd.sample(samples_by_k[k], max_samples)
        path = os.path.join(output_dir, f"object_perception_{dimension_name}_k{k}_{split}_{max_samples}.jsonl")
        with open(path, "w", encoding="utf-8") as f:
            for s in samples_by_k[k]:
                f.write(json.dumps(s) + "\n")
        print(f"Written K={k} {len(samples_by_k[k])} samples to {path}")
    print(f"Finished building QA samples for {dimension_name}.")
    return samples_by_k


def convert_train_sample_to_eval_sample(train_sample):
    conversation = train_sample.pop("conversations")
    train_sample["text"] = conversation[0]["value"]
    return train_sample


def build_train_and_val_datasets(scene_info_path="data/scannet/scannet_instance_data/scenes_train_val_info_i_D5.pkl",
                                 train_dir="training_data/object_perception", val_dir="evaluation_data/object_perception",
                                 max_train_samples=-1, val_max_samples=3000, templates=None, rnd=random):
    """SOPE:229-268: train files as they are, val files built in train format under ``temp`` then converted."""
    os.makedirs(train_dir, exist_ok=True)
    os.makedirs(val_dir, exist_ok=True)
    h = _handler(scene_info_path)
    for dim in ("height", "length", "width"):
        build_lwh_qa_samples(h, os.path.join(train_dir, f"merged_train_object_coverage_{dim}.pkl"), dim, "train", train_dir,
                             max_k=6, max_samples=max_train_samples, templates=templates, rnd=rnd)
    temp = os.path.join(val_dir, "temp")
    os.makedirs(temp, exist_ok=True)
    for dim in ("height", "length", "width"):
        build_lwh_qa_samples(h, os.path.join(val_dir, f"merged_val_object_coverage_{dim}.pkl"), dim, "val", temp,
                             max_k=6, max_samples=val_max_samples, templates=templates, rnd=rnd)
    for fname in os.listdir(temp):
        with open(os.path.join(temp, fname), encoding="utf-8") as fin, open(os.path.join(val_dir, fname), "w", encoding="utf-8") as fout:
            for line in fin:
                fout.write(json.dumps(convert_train_sample_to_eval_sample(json.loads(line))) + "\n")
    shutil.rmtree(temp)
    print("Finished building both TRAIN and VAL datasets.")


def main(argv=None):
    """``python -m multi_spatialmllm_b200.object_perception [--reference_root ...]`` = SOPE's ``main()`` (seed 1, SOPE:14)."""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene_info_path", default="data/scannet/scannet_instance_data/scenes_train_val_info_i_D5.pkl")
    ap.add_argument("--reference_root", default=None)
    a = ap.parse_args(argv)
    random.seed(1)
    build_train_and_val_datasets(a.scene_info_path, templates=load_reference_templates(a.reference_root) if a.reference_root else None)


if __name__ == "__main__":
    main()
