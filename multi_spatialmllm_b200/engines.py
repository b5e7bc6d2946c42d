"""Drop-in QA engines: the reference's per-task generators with the geometry on the GPU.

Each class keeps the name, constructor arguments, methods and output files of its reference counterpart and consumes
Python's ``random`` stream call for call, so that with the same seed it writes the same JSONL the reference writes —
while every per-sample ``get_point_2d_coordinates_in_image`` call (IH:291-305: a full ``np.load`` of the scene, two JPEG
decodes and a PNG decode per call in the reference) becomes an array lookup in a per-frame table produced by ONE
``msp_annotate_points`` launch, and the visibility lists (JSON strings in the reference) are views of the bit masks K1 wrote.

  DepthEstimationCoorQAEngine   <- depth_perception/depth_estimation_coor_engine.py:35-312
  DepthEstimationDotQAEngine    <- depth_perception/depth_estimation_dot_engine.py:35-323   (+ annotated JPEGs on the GPU)
  DepthComparisonCoorQAEngine   <- depth_perception/depth_comparison_coor_engine.py:36-402
  DepthComparisonDotQAEngine    <- depth_perception/depth_comparison_dot_engine.py:36-433   (+ annotated JPEGs on the GPU)

Phrasings are not vendored (CC BY-NC): pass ``templates={"task": [...], "questions": [...], ...}`` or load them from a
reference checkout with ``load_reference_templates``; the neutral defaults of ``streams.Templates`` are used otherwise.
The visibility table is recomputed on the device (it is bit-identical to the reference's parquet, tests/test_gpu_parity.py),
so ``visibility_info_path`` is accepted for signature compatibility and not read.
"""
from __future__ import annotations

import collections.abc
import json
import os
import random

import numpy as np

from . import streams, tables
from .handler.info_handler import SceneInfoHandler


# ------------------------------------------------------------------------------------------------
# scene residency + per-frame lookup tables
# ------------------------------------------------------------------------------------------------
class _IntSeq(collections.abc.Sequence):
    """A sorted id array that random.sample / random.choices / len() can use like the reference's Python list; items
    come out as Python ints (what json.loads gave the reference)."""

    def __init__(self, arr):
        self.arr = arr

    def __len__(self):
        return len(self.arr)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return _IntSeq(self.arr[i])
        return int(self.arr[i])


class SceneSession:
    """One scene resident in HBM for the engines: points, poses, depth maps, K1's masks; per-frame annotation tables."""

    def __init__(self, scene_info: SceneInfoHandler, scene_id: str, device="cuda:0"):
        self.scene_id = scene_id
        self.sc, self.image_ids = tables.device_scene_from_handler(scene_info, scene_id, device)
        self.sc.project_visibility()
        self.row = {i: f for f, i in enumerate(self.image_ids)}
        self.mask_words = self.sc.mask.cpu().numpy().view(np.uint32)          # (F, stride): 25 KB per frame
        self._visible, self._table = {}, {}

    def visible_points(self, image_id) -> _IntSeq:
        """get_image_to_points_info (IH:502-519): ascending ids of the points visible in the frame."""
        if image_id not in self._visible:
            bits = np.unpackbits(self.mask_words[self.row[image_id]].view(np.uint8), bitorder="little")[:self.sc.n]
            self._visible[image_id] = np.flatnonzero(bits)
        return _IntSeq(self._visible[image_id])

    def lookup(self, image_id, point: int):
        """get_point_2d_coordinates_in_image(check_visible=True, return_depth=True) + the integer maths every engine
        applies to it (DE-dot:216-218, 261): (visible, x_permille, y_permille, depth_mm, (int(u), int(v))).
        All visible points of the frame are annotated by one launch the first time the frame is touched."""
        if image_id not in self._table:
            ids = self.visible_points(image_id).arr
            f = self.row[image_id]
            a = self.sc.annotate_points(ids.astype(np.int32), np.full(len(ids), f, np.int32))
            self._table[image_id] = {k: v.cpu().numpy() for k, v in a.items()}
        ids, t = self._visible[image_id], self._table[image_id]
        k = int(np.searchsorted(ids, point))
        if k >= len(ids) or ids[k] != point:                                    # not in the table: a direct request
            a = self.sc.annotate_points(np.array([point], np.int32), np.array([self.row[image_id]], np.int32))
            t, k = {key: v.cpu().numpy() for key, v in a.items()}, 0
        return (bool(t["visible"][k]), int(t["permille"][k][0]), int(t["permille"][k][1]), int(t["depth_mm"][k]),
                (int(t["pix"][k][0]), int(t["pix"][k][1])))


def load_reference_templates(reference_root: str, relative_path: str, attrs=("task_description", "templates")) -> dict:
    """The ``self.task_description`` / ``self.templates`` literals an engine class assigns in ``__init__``, read from a
    reference checkout through the AST (nothing is imported or executed).  Returns {"task": [...], **templates}."""
    import ast
    tree = ast.parse(open(os.path.join(reference_root, relative_path)).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1:
            t = node.targets[0]
            if isinstance(t, ast.Attribute) and isinstance(t.value, ast.Name) and t.value.id == "self" and t.attr in attrs:
                try:
                    found[t.attr] = ast.literal_eval(node.value)
                except ValueError:
                    pass
    if set(attrs) - set(found):
        raise KeyError(f"{relative_path}: no literal self.{sorted(set(attrs) - set(found))}")
    return {"task": found["task_description"], **found["templates"]}


def _distinct_colors(rnd, n, max_retries=10):
    """generate_distinct_colors (DE-dot:22-33) on ``rnd``."""
    colors, retries = [], 0
    while len(colors) < n and retries < max_retries:
        c = (rnd.randint(0, 255), rnd.randint(0, 255), rnd.randint(0, 255))
        if all(sum(abs(a - b) for a, b in zip(c, e)) > 300 for e in colors):
            colors.append(c)
        retries += 1
    if len(colors) < n:
        colors += rnd.sample([(255, 0, 0), (0, 255, 0), (0, 0, 255), (0, 0, 0), (255, 255, 255)], n - len(colors))
    return colors


# ------------------------------------------------------------------------------------------------
# shared driver (generate_qa_training_data / generate_qa_eval_data are the same in the four depth engines)
# ------------------------------------------------------------------------------------------------
class _DepthEngine:
    task_name = ""
    template_key = ""
    reference_file = ""

    def __init__(self, scene_info_path, version_num="v1_0", all_max_samples=-1, image_output_dir=None,
                 visibility_info_path=None, max_n_points_per_image=1, warning_file=None, *,
                 posed_images_root="data/scannet/posed_images", instance_data_root="data/scannet/scannet_instance_data",
                 templates=None, rnd=random, device="cuda:0"):
        self.scene_info = scene_info_path if isinstance(scene_info_path, SceneInfoHandler) else SceneInfoHandler(
            scene_info_path, posed_images_root=posed_images_root, instance_data_root=instance_data_root)
        self.version_num, self.image_output_dir, self.all_max_samples = version_num, image_output_dir, all_max_samples
        self.max_n_points_per_image, self.warning_file = max_n_points_per_image, warning_file
        self.visibility_info_path = visibility_info_path
        t = templates or streams.Templates.NEUTRAL[self.template_key]
        self.task_description = t["task"]
        self.templates = {k: v for k, v in t.items() if k != "task"}
        self.rnd, self.device = rnd, device
        self._session = None
        self.pending_marks = []            # (scene_id, image_id, output name, [marks]) of the dot engines, drained per scene

    # -- plumbing
    def _warn(self, message):
        print(message.strip())
        if self.warning_file:
            with open(self.warning_file, "a") as wf:
                wf.write(message.strip())

    def session(self, scene_id) -> SceneSession:
        if self._session is None or self._session.scene_id != scene_id:
            self._session = SceneSession(self.scene_info, scene_id, self.device)
        return self._session

    def _flush_images(self, scene_id):
        """Paint and write the annotated JPEGs queued for ``scene_id`` (cv2.imread / circle / putText / imwrite of the dot
        engines) on the device: annotate.write_annotated (nvJPEG decode once per frame, msp_draw_marks, nvJPEG encode)."""
        if not self.pending_marks:
            return
        from . import annotate
        jobs, self.pending_marks = self.pending_marks, []
        annotate.write_annotated(jobs, lambda image_id: open(self.scene_info.get_image_path(scene_id, image_id), "rb").read(),
                                 os.path.join(self.image_output_dir, scene_id), self.device)

    # -- the reference's split drivers (DE-coor:256-312; identical in the other three files)
    def generate_qa_training_data(self, output_dir, save_file=True):
        scene_ids = self.scene_info.get_sorted_keys()
        if self.all_max_samples > 0:
            self.max_samples = max(self.all_max_samples // len(scene_ids) + 1, 1)
            if self.max_samples == 1:
                scene_ids = self.rnd.sample(scene_ids, self.all_max_samples)
        else:
            self.max_samples = -1
        self.num_used_scenes = len(scene_ids)
        train_data = []
        for scene_id in scene_ids:
            train_data.extend(self.generate_qa_training_single_scene(scene_id))
            self._flush_images(scene_id)
        if len(train_data) > self.all_max_samples:
            train_data = self.rnd.sample(train_data, self.all_max_samples)
        self.rnd.shuffle(train_data)
        if not save_file:
            return train_data
        os.makedirs(output_dir, exist_ok=True)
        path = f"{output_dir}/{self.task_name}.jsonl"
        with open(path, "w") as f:
            for entry in train_data:
                f.write(json.dumps(entry) + "\n")
        print(f"[Train] Training data saved to {path}. Generated {len(train_data)} samples in total.")

    @staticmethod
    def convert_train_sample_to_eval_sample(train_sample):
        train_sample["text"] = train_sample["conversations"][0]["value"]
        return train_sample

    def generate_qa_eval_data(self, output_dir):
        assert self.max_n_points_per_image == 1, "max_n_points_per_image should be 1 for evaluation"
        data = [self.convert_train_sample_to_eval_sample(s) for s in self.generate_qa_training_data(output_dir, save_file=False)]
        os.makedirs(output_dir, exist_ok=True)
        path = f"{output_dir}/{self.task_name}.jsonl"
        with open(path, "w") as f:
            for entry in data:
                f.write(json.dumps(entry) + "\n")
        print(f"[Eval] Evaluation data saved to {path}. Generated {len(data)} samples in total.")

    def _sample_images(self, image_ids, with_replacement_when_short):
        """DE-coor:181-188 / DC-coor:246-257."""
        if self.max_samples > 0:
            if with_replacement_when_short and self.max_samples > len(image_ids):
                return self.rnd.choices(image_ids, k=self.max_samples)
            return self.rnd.sample(image_ids, min(self.max_samples, len(image_ids)))
        return self.rnd.sample(image_ids, len(image_ids))


class _DepthEstimation(_DepthEngine):
    dot = False

    def generate_qa_training_single_scene(self, scene_id):
        """DE-coor:175-254 / DE-dot:175-265."""
        ses, rnd = self.session(scene_id), self.rnd
        image_ids = self.scene_info.get_all_extrinsic_valid_image_ids(scene_id)
        H, W = self.scene_info.get_image_shape(scene_id)
        out = []
        for image_id in self._sample_images(image_ids, False):
            visible_points = ses.visible_points(image_id)
            k = self.max_n_points_per_image
            sampled = rnd.choices(visible_points, k=k) if len(visible_points) < k else rnd.sample(visible_points, k)
            for point in sampled:
                vis, x, y, depth, (u, v) = ses.lookup(image_id, point)
                if not vis:
                    self._warn(f"Warning: Point-Id {point} is not visible in image {image_id} in scene {scene_id}.\n")
                    continue
                if self.dot:                                                    # DE-dot:220-229 (colour before the templates)
                    color = _distinct_colors(rnd, 1)[0]
                    name = f"{image_id}_p{point}_annotated.jpg"
                    self.pending_marks.append((image_id, name, [(u, v, 10, -1, 0, 0, color, 0)]))
                q = rnd.choice(self.templates["questions"])
                question = q if self.dot else q.format(x1=x, y1=y)
                answer = rnd.choice(self.templates["answers"]).format(x1=x, y1=y, depth=depth)
                task = rnd.choice(self.task_description)
                out.append({"id": f"{scene_id}_{image_id}_point{point}",
                            "image": [f"{scene_id}/{image_id}_p{point}_annotated.jpg" if self.dot else f"{scene_id}/{image_id}.jpg"],
                            "conversations": [{"from": "human", "value": f"{task}\n{question}"}, {"from": "gpt", "value": answer}],
                            "height_list": [H], "width_list": [W], "question_type": self.task_name, "gt_value": depth,
                            "ori_coordinates": [u, v]})
        return out


class DepthEstimationCoorQAEngine(_DepthEstimation):
    task_name, template_key = "depth_estimation_coor", "depth_estimation"
    reference_file = "spatial_engine/depth_perception/depth_estimation_coor_engine.py"


class DepthEstimationDotQAEngine(_DepthEstimation):
    task_name, template_key, dot = "depth_estimation_dot", "depth_estimation_dot", True
    reference_file = "spatial_engine/depth_perception/depth_estimation_dot_engine.py"


class _DepthComparison(_DepthEngine):
    dot = False

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        assert self.max_n_points_per_image == 1, "single-round QAs only (DC-coor:54-55)"

    def _pair_info(self, ses, scene_id, image_id, points_pair):
        info = []
        for i, p in enumerate(points_pair):
            vis, x, y, depth, coords = ses.lookup(image_id, p)
            if not vis:
                self._warn(f"Warning: Point-Id {p} is not visible in image {image_id} in scene {scene_id}.\n")
                continue
            info.append({"x": x, "y": y, "depth": depth, "coords": coords, "letter": chr(65 + i)})
        return info

    def generate_qa_training_single_scene(self, scene_id):
        """DC-coor:240-346 (one attempt per pair) / DC-dot:240-375 (up to 11 attempts, letters and dots on the image)."""
        ses, rnd = self.session(scene_id), self.rnd
        image_ids = self.scene_info.get_all_extrinsic_valid_image_ids(scene_id)
        H, W = self.scene_info.get_image_shape(scene_id)
        if self.dot:
            from .annotate import default_atlas
            atlas = default_atlas()
        out = []
        for image_id in self._sample_images(image_ids, True):
            visible_points = ses.visible_points(image_id)
            for _ in range(self.max_n_points_per_image):
                retry = 0
                while True:
                    points_pair = rnd.sample(visible_points, 2)
                    info = self._pair_info(ses, scene_id, image_id, points_pair)
                    if len(info) == 2 and info[0]["depth"] != info[1]["depth"]:
                        break
                    if len(info) == 2:
                        self._warn(f"Warning: Points {points_pair} in image {image_id} in scene {scene_id} have the same depth.\n Skip this pair.")
                    else:
                        self._warn(f"Warning: Cannot find two visible points in image {image_id} in scene {scene_id}.\n Skip this pair.")
                    retry += 1
                    if not self.dot or retry > 10:
                        info = None
                        break
                if info is None:
                    if self.dot:
                        self._warn(f"Failed to find valid pair after 10 retries in image {image_id} in scene {scene_id}.")
                    continue
                letters = ["A", "B"]
                rnd.shuffle(letters)
                shuffled = rnd.sample(info, 2)
                for i, pi in enumerate(shuffled):
                    pi["letter"] = letters[i]
                p1, p2 = shuffled
                closer, farther = (p1, p2) if p1["depth"] <= p2["depth"] else (p2, p1)
                is_closer = rnd.choice([True, False])
                kind = "closer" if is_closer else "farther"
                q = rnd.choice(self.templates[f"{kind}_questions"])
                a = rnd.choice(self.templates[f"{kind}_answers"])
                task = rnd.choice(self.task_description)
                pick = closer if is_closer else farther
                stem = f"{image_id}_p{points_pair[0]}_p{points_pair[1]}"
                if self.dot:                                                    # DC-dot:333-346
                    marks = []
                    for pi in shuffled:
                        color = (rnd.randint(0, 255), rnd.randint(0, 255), rnd.randint(0, 255))
                        cx, cy = pi["coords"]
                        marks.append((cx, cy, 10, atlas.index(pi["letter"]), cx + 15, cy + 15, color, 0))
                    self.pending_marks.append((image_id, f"{stem}_annotated.jpg", marks))
                    question, answer, gt = q, a.format(correct_label=pick["letter"]), pick["letter"]
                    qtype, image = "depth_comparison_annotated", f"{scene_id}/{stem}_annotated.jpg"
                else:
                    question = q.format(x1=p1["x"], y1=p1["y"], x2=p2["x"], y2=p2["y"])
                    answer, gt = a.format(correct_x=pick["x"], correct_y=pick["y"]), [pick["x"], pick["y"]]
                    qtype, image = "depth_comparison_coordinate", f"{scene_id}/{image_id}.jpg"
                out.append({"id": f"{scene_id}_{stem}", "image": [image],
                            "conversations": [{"from": "human", "value": f"{task}\n{question}"}, {"from": "gpt", "value": answer}],
                            "height_list": [H], "width_list": [W], "question_type": qtype, "gt_value": gt,
                            "points_info": shuffled, "is_closer_question": is_closer})
        return out


class DepthComparisonCoorQAEngine(_DepthComparison):
    task_name, template_key = "depth_comparison_coor", "depth_comparison"
    reference_file = "spatial_engine/depth_perception/depth_comparison_coor_engine.py"


class DepthComparisonDotQAEngine(_DepthComparison):
    task_name, template_key, dot = "depth_comparison_dot", "depth_comparison_dot", True
    reference_file = "spatial_engine/depth_perception/depth_comparison_dot_engine.py"


def main(argv=None):
    """The ``__main__`` of the four depth engines (DE-dot:319-370 and siblings): same flags, same output layout."""
    import argparse
    engines = {c.task_name: c for c in (DepthEstimationCoorQAEngine, DepthEstimationDotQAEngine, DepthComparisonCoorQAEngine,
                                        DepthComparisonDotQAEngine)}
    ap = argparse.ArgumentParser()
    ap.add_argument("task", choices=sorted(engines))
    ap.add_argument("--output_suffix", default="")
    ap.add_argument("--train_scene_info_path", default="data/scannet/scannet_instance_data/scenes_train_info_i_D5.pkl")
    ap.add_argument("--val_scene_info_path", default="data/scannet/scannet_instance_data/scenes_val_info_i_D5.pkl")
    ap.add_argument("--train_all_max_samples", type=int, default=500000)
    ap.add_argument("--val_all_max_samples", type=int, default=300)
    ap.add_argument("--output_dir_train", default=None)
    ap.add_argument("--output_dir_val", default=None)
    ap.add_argument("--version_num", default="v1_0")
    ap.add_argument("--reference_root", default=None, help="reference checkout to read the phrasings from")
    ap.add_argument("--seed", type=int, default=None, help="default: the seed the reference module sets (4/5/6)")
    a = ap.parse_args(argv)
    cls = engines[a.task]
    out_train = os.path.join(a.output_dir_train or f"training_data/{a.task}", a.version_num, a.output_suffix.replace("_", ""))
    out_val = os.path.join(a.output_dir_val or f"evaluation_data/{a.task}", a.version_num, a.output_suffix.replace("_", ""))
    tp = load_reference_templates(a.reference_root, cls.reference_file) if a.reference_root else None
    seed = a.seed if a.seed is not None else {"depth_estimation_coor": 4, "depth_estimation_dot": 5}.get(a.task, 6)
    random.seed(seed)
    np.random.seed(seed)
    for info, n, out in ((a.val_scene_info_path, a.val_all_max_samples, out_val), (a.train_scene_info_path, a.train_all_max_samples, out_train)):
        os.makedirs(out, exist_ok=True)
        eng = cls(info, a.version_num, n, os.path.join(out, "images"), None, warning_file=os.path.join(out, "warning.txt"), templates=tp)
        (eng.generate_qa_eval_data if out is out_val else eng.generate_qa_training_data)(out)


if __name__ == "__main__":
    main()
