"""Drop-in pair-table QA engines: camera_movement and the two visual_correspondence generators, geometry on the GPU.

Reference counterparts (module-level functions with a ``main()``; the public names are kept as methods / functions):

  sample_dataframe                    <- camera_movement/camera_movement_engine_train_val.py:29-151 (copied verbatim into
                                         VC-coor:70-192 and VC-dot:138-260 upstream; one implementation here)
  CameraMovementEngine                <- camera_movement/camera_movement_engine_train_val.py:153-446
  VisualCorrespondenceCoorEngine      <- visual_correspondence/visual_correspondence_qa_engine_coor_2_coor.py:264-540
  VisualCorrespondenceDotEngine       <- visual_correspondence/visual_correspondence_qa_engine_dot_2_multichoice.py:280-577

The reference walks the sampled pair table row by row and, per row, intersects two JSON-decoded visibility lists
(``np.intersect1d`` on ~10^4-10^5 ints), reloads the scene's point file twice and decodes four JPEGs and two PNGs
(IH:284-305) — 4 h per million rows by its own remark (VC-coor:526).  Here the rows are grouped by scene; per scene ONE
``msp_pair_overlap`` launch counts the common visible points of every row's frame pair straight from K1's bit masks and
ONE ``msp_pair_annotate`` launch returns, per row, the k-th common point re-projected into both frames (per-mille and
pixel integers, visibility re-check) together with ``inv(E1) @ E2``.  The host only replays Python's ``random`` stream:

* the draws of a row depend on the device results in two rare ways (a failed visibility re-check skips the phrasing
  draws; a random distractor that hits the correct pixel is redrawn, VC-dot:368-373), so the stream is replayed
  speculatively: a dry run over the remaining rows with the common case predicted yields every row's request
  (ordered frame pair, index of the drawn common point), the device answers them scene by scene, and the real run
  consumes the answers; a row whose answer contradicts the prediction finishes on its real answer and the rows after
  it are re-planned from the RNG position it left.  The output is the reference's, draw for draw.
* ``DataFrame.sample`` (global NumPy RNG) is replayed by ``np.random.choice(len, n, replace=False)``, the call pandas
  makes (pandas/core/sample.py), on groups of the same sizes in the same order.

Phrasings are not vendored (CC BY-NC): pass ``templates=`` or read them from a reference checkout with
``load_reference_templates``; ``streams.Templates.NEUTRAL`` is the fall-back.
"""
from __future__ import annotations

import json
import os
import random

import numpy as np
import pandas as pd

from . import streams, tables
from .handler.info_handler import SceneInfoHandler
from .streams import CAMERA_QUESTION_TYPES, camera_answer_values


# ------------------------------------------------------------------------------------------------
# overlap-bin sampler
# ------------------------------------------------------------------------------------------------
def _pandas_groupby_observed_default() -> bool:
    """pandas >= 3 iterates observed categories only; pandas 2.x (the reference pins 2.2.3, data_engine.yaml:196)
    also yields the empty bins, which take part in the quota split (CME:86-93)."""
    return int(pd.__version__.split(".")[0]) >= 3


def sample_rows_by_overlap(overlap, all_overlap_samples, non_overlap_samples, overlap_min=0, overlap_max=100, interval=1,
                           observed=None, np_random=np.random, log=print) -> np.ndarray:
    """Row numbers ``sample_dataframe`` (CME:29-151) selects, in its output order, consuming ``np_random`` like it does.

    Bins are ``pd.cut(overlap, arange(min, max + interval, interval), include_lowest=True)``: right-closed, the first
    one also closed on the left; NaN and out-of-range rows are dropped.  The quota is split evenly over the bins in bin
    order (first ``remainder`` bins get one more), the bins are then visited from the smallest to the largest (stable),
    a bin smaller than its quota is taken whole and hands the rest of its quota to the next one."""
    ov = np.asarray(overlap, dtype=np.float64)
    observed = _pandas_groupby_observed_default() if observed is None else observed
    zero = np.flatnonzero(ov == 0)
    if len(zero) > non_overlap_samples:                                           # DataFrame.sample(n=...)
        zero = zero[np_random.choice(len(zero), size=non_overlap_samples, replace=False)]
    edges = np.arange(overlap_min, overlap_max + interval, interval)
    rest = np.flatnonzero(ov != 0)                                                # NaN != 0: dropped below, as by dropna
    x = ov[rest]
    which = np.searchsorted(edges, x, side="left")
    which[x == edges[0]] = 1
    keep = ~np.isnan(x) & (which > 0) & (which < len(edges))
    rest, which = rest[keep], which[keep] - 1
    n_bins = len(edges) - 1
    order = np.argsort(which, kind="stable")
    bounds = np.searchsorted(which[order], np.arange(n_bins + 1))
    groups = [rest[order[bounds[b]:bounds[b + 1]]] for b in range(n_bins)]
    if observed:
        groups = [g for g in groups if len(g)]
    elif len(rest) == 0:
        groups = []
    if not groups:
        return zero
    base, remainder = divmod(all_overlap_samples, len(groups))
    quotas = [base + (1 if b < remainder else 0) for b in range(len(groups))]
    taken, leftover = [], 0
    for b in sorted(range(len(groups)), key=lambda b: len(groups[b])):
        g, quota = groups[b], quotas[b] + leftover
        log(f"[sample_dataframe] current_quota v.s. group_df: {quota} v.s. {len(g)}")
        if len(g) <= quota:
            taken.append(g)
            leftover = quota - len(g)
        else:
            taken.append(g[np_random.choice(len(g), size=quota, replace=False)])
            leftover = 0
    if leftover > 0:
        log(f"[sample_dataframe] Warning: bins not enough to reach {all_overlap_samples}; leftover {leftover}")
    return np.concatenate(taken + [zero])


def sample_dataframe(df, all_overlap_samples, non_overlap_samples, overlap_min=0, overlap_max=100, interval=1, observed=None):
    """The reference's signature: the sampled rows as a DataFrame with a fresh RangeIndex."""
    idx = sample_rows_by_overlap(df["overlap"].to_numpy(), all_overlap_samples, non_overlap_samples, overlap_min,
                                 overlap_max, interval, observed)
    return df.iloc[idx].reset_index(drop=True)


# ------------------------------------------------------------------------------------------------
# scene residency
# ------------------------------------------------------------------------------------------------
class SceneCache:
    """Scenes resident in HBM (points, poses, depth maps, K1's masks), least recently used first out."""

    def __init__(self, scene_infos: SceneInfoHandler, device="cuda:0", budget_bytes=96 << 30):
        self.scene_infos, self.device, self.budget = scene_infos, device, budget_bytes
        self._scenes, self._bytes = {}, {}

    def get(self, scene_id):
        """(DeviceScene, {image_id: row})"""
        if scene_id in self._scenes:
            self._scenes[scene_id] = self._scenes.pop(scene_id)                  # most recently used last
            return self._scenes[scene_id]
        sc, image_ids = tables.device_scene_from_handler(self.scene_infos, scene_id, self.device)
        sc.project_visibility()
        size = sum(t.numel() * t.element_size() for t in (sc.pts, sc.depth, sc.mask) if t is not None)
        while self._scenes and sum(self._bytes.values()) + size > self.budget:
            old = next(iter(self._scenes))
            del self._scenes[old], self._bytes[old]
        self._scenes[scene_id], self._bytes[scene_id] = (sc, {i: f for f, i in enumerate(image_ids)}), size
        return self._scenes[scene_id]


def _by_scene(scene_ids):
    """[(scene_id, row numbers)] with each scene's rows together (table order inside a scene)."""
    scene_ids = np.asarray(scene_ids, dtype=object)
    if len(scene_ids) == 0:
        return []
    uniq, inv = np.unique(scene_ids.astype(str), return_inverse=True)
    order = np.argsort(inv, kind="stable")
    cuts = np.searchsorted(inv[order], np.arange(len(uniq) + 1))
    return [(str(uniq[k]), order[cuts[k]:cuts[k + 1]]) for k in range(len(uniq))]


def load_reference_templates(reference_root: str, engine: str) -> dict:
    """Phrasings of ``engine`` in {"camera_movement", "visual_correspondence", "visual_correspondence_dot"} from a
    reference checkout, read through the AST (``streams.Templates._literals``: nothing is imported or executed)."""
    lit = streams.Templates._literals
    if engine == "camera_movement":
        t = lit(os.path.join(reference_root, "spatial_engine/camera_movement/TEMPLATES.py"),
                ("QUESTION_TEMPLATES", "ANSWER_TEMPLATES", "TASK_DESCRIPTION"))
        return {"task": t["TASK_DESCRIPTION"], "questions": t["QUESTION_TEMPLATES"], "answers": t["ANSWER_TEMPLATES"]}
    rel = {"visual_correspondence": "visual_correspondence_qa_engine_coor_2_coor.py",
           "visual_correspondence_dot": "visual_correspondence_qa_engine_dot_2_multichoice.py"}[engine]
    t = lit(os.path.join(reference_root, "spatial_engine/visual_correspondence", rel), ("TASK_DESCRIPTION", "TEMPLATES"))
    return {"task": t["TASK_DESCRIPTION"], **t["TEMPLATES"]}


# ------------------------------------------------------------------------------------------------
# shared driver: plan (dry run) -> device -> real run
# ------------------------------------------------------------------------------------------------
class _PairTableEngine:
    template_key = ""
    needs_common_point = True

    def __init__(self, scene_infos, templates=None, rnd=random, device="cuda:0", cache: SceneCache | None = None,
                 posed_images_root="data/scannet/posed_images", instance_data_root="data/scannet/scannet_instance_data"):
        self.scene_infos = scene_infos if isinstance(scene_infos, SceneInfoHandler) else SceneInfoHandler(
            scene_infos, posed_images_root=posed_images_root, instance_data_root=instance_data_root)
        t = templates or streams.Templates.NEUTRAL[self.template_key]
        self.task_description = t["task"]
        self.templates = {k: v for k, v in t.items() if k != "task"}
        self.rnd, self.device = rnd, device
        self.cache = cache or SceneCache(self.scene_infos, device)
        self.replans = 0                                   # times a device answer contradicted the prediction

    # -- per-row logic, supplied by the subclasses -------------------------------------------------
    def _request(self, ctx, r):
        """First draws of row r: the ordered pair and the indices of the drawn common points -> (swap, [k, ...]) or
        None when the row is dropped before any device result is needed."""
        raise NotImplementedError

    def _finish(self, ctx, r, req, recs, dry):
        """Remaining draws of row r.  ``recs`` is None in the dry run (common case predicted).  -> (sample | None,
        whether the device answer matched the prediction)."""
        raise NotImplementedError

    # -- plumbing ----------------------------------------------------------------------------------
    def _warn(self, ctx, message):
        print(message.strip())
        if ctx["warning_file"]:
            with open(ctx["warning_file"], "a") as wf:
                wf.write(message)

    def _context(self, df, warning_file=None, **extra):
        ctx = {"scene": df["scene_id"].tolist(), "i1": df["image_id1"].tolist(), "i2": df["image_id2"].tolist(),
               "overlap": df["overlap"].to_numpy(np.float64), "yaw": df["yaw"].to_numpy(np.float64),
               "pitch": df["pitch"].to_numpy(np.float64), "distance": df["distance"].to_numpy(np.float64),
               "warning_file": warning_file, "dry": False, **extra}
        R = len(df)
        ctx["inter"] = np.zeros(R, np.int64)
        if self.needs_common_point:                          # |visible(i1) & visible(i2)| of every row: one launch per scene
            for scene_id, rows in _by_scene(ctx["scene"]):
                sc, row_of = self.cache.get(scene_id)
                pairs = np.array([(row_of[ctx["i1"][r]], row_of[ctx["i2"][r]]) for r in rows], np.int32)
                ctx["inter"][rows] = sc.pair_overlap(pairs)[0].cpu().numpy()
        return ctx

    def _device_answers(self, ctx, plan, first):
        """One msp_pair_annotate launch per scene over every request planned for the rows >= first."""
        slots = [(r, q) for r, req in enumerate(plan, start=first) if req is not None for q in range(len(req[1]))]
        answers = [None if req is None else [None] * len(req[1]) for req in plan]
        if not slots:
            return answers
        slot_rows = np.array([r for r, _ in slots])
        for scene_id, which in _by_scene([ctx["scene"][r] for r in slot_rows]):
            sc, row_of = self.cache.get(scene_id)
            pairs, sel = np.empty((len(which), 2), np.int32), np.empty(len(which), np.uint32)
            for n, w in enumerate(which):
                r, q = slots[w]
                swap, ks = plan[r - first]
                a, b = row_of[ctx["i1"][r]], row_of[ctx["i2"][r]]
                pairs[n] = (b, a) if swap else (a, b)
                sel[n] = ks[q]
            rec = sc.records_to_numpy(sc.pair_annotate(pairs, sel, sel_mode=0))
            for n, w in enumerate(which):
                r, q = slots[w]
                answers[r - first][q] = rec[n]
        return answers

    def build_samples(self, df, warning_file=None, plan_rows=1 << 18, **extra):
        """The row loop of build_train_dataset / build_val_dataset: one entry per row of ``df`` (None = row dropped).
        At most ``plan_rows`` rows are planned ahead per round trip to the device."""
        ctx = self._context(df, warning_file, **extra)
        R, out, first = len(df), [], 0
        while first < R:
            state = self.rnd.getstate()
            ctx["dry"] = True
            plan, last = [], min(R, first + plan_rows)
            for r in range(first, last):
                req = self._request(ctx, r)
                plan.append(req)
                if req is not None:
                    self._finish(ctx, r, req, None, True)
            self.rnd.setstate(state)
            ctx["dry"] = False
            answers = self._device_answers(ctx, plan, first)
            for r in range(first, last):
                req = self._request(ctx, r)
                assert req == plan[r - first], "RNG replay diverged from its own plan"
                if req is None:
                    out.append(None)
                    continue
                sample, as_predicted = self._finish(ctx, r, req, answers[r - first], False)
                out.append(sample)
                if not as_predicted:                         # the rows after r were planned on a wrong RNG position
                    self.replans += 1
                    break
            first = len(out)
        return out

    @staticmethod
    def convert_train_sample_to_eval_sample(train_sample):
        conversation = train_sample.pop("conversations")
        train_sample["text"] = conversation[0]["value"]
        return train_sample

    def _write(self, samples, path):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "w") as f:
            for item in samples:
                f.write(json.dumps(item) + "\n")


# ------------------------------------------------------------------------------------------------
# camera movement
# ------------------------------------------------------------------------------------------------
class CameraMovementEngine(_PairTableEngine):
    """CME:153-245 per row; CME:269-350 per split and question type."""
    template_key, needs_common_point = "camera_movement", False

    def _request(self, ctx, r):
        return (self.rnd.random() < 0.5, [0])                                    # CME:163

    def _finish(self, ctx, r, req, recs, dry):
        rnd, qtype = self.rnd, ctx["qtype"]
        task = rnd.choice(self.task_description)
        if float(ctx["overlap"][r]) < 0.1:                                       # CME:198-201
            raise NotImplementedError("overlap < 0.1 is not supported yet.")
        question = rnd.choice(self.templates["questions"][qtype])
        answer_template = rnd.choice(self.templates["answers"][qtype])
        if dry:
            return None, True
        swap, rec = req[0], recs[0]
        scene_id, image1, image2 = ctx["scene"][r], ctx["i1"][r], ctx["i2"][r]
        yaw, pitch = float(ctx["yaw"][r]), float(ctx["pitch"][r])
        if swap:                                                                 # CME:163-166: the table's angles, negated
            yaw, pitch, image1, image2 = -yaw, -pitch, image2, image1
        assert not np.isnan(rec["rel_disp"]).any(), f"pose is nan for {scene_id} {image1} {image2}"   # CME:182-183
        av = camera_answer_values(rec["rel_disp"], yaw, pitch)                   # CME:168-172, 205-225
        assert abs(float(np.linalg.norm(rec["rel_disp"])) - ctx["distance"][r]) < 0.1, \
            f"distance is not close to the distance from df for {scene_id} {image1} {image2}."        # CME:193
        H, W = self.scene_infos.get_image_shape(scene_id, image1)
        return {"id": ctx["idx0"] + r, "image": [f"{scene_id}/{image1}.jpg", f"{scene_id}/{image2}.jpg"],
                "conversations": [{"from": "human", "value": f"{task}\n{question}"},
                                  {"from": "gpt", "value": answer_template.format(**av)}],
                "height_list": [H] * 2, "width_list": [W] * 2, "answer_values": av, "question_type": qtype,
                "gt_value": av[qtype]}, True

    def _build(self, parquet_path, output_dir, qtype, desired_count, overlap_min, overlap_max, interval, split):
        assert qtype in CAMERA_QUESTION_TYPES
        df = pd.read_parquet(parquet_path)
        tag = "Train" if split == "train" else "Val"
        print(f"[{tag}: {qtype}] Loaded DF with {len(df)} rows from {parquet_path}")
        df_sampled = sample_dataframe(df, desired_count, 0, overlap_min, overlap_max, interval)
        print(f"[{tag}: {qtype}] got {len(df_sampled)} sampled rows")
        samples = self.build_samples(df_sampled, qtype=qtype, idx0=0)
        if split == "val":
            samples = [self.convert_train_sample_to_eval_sample(s) for s in samples]
        self.rnd.shuffle(samples)
        out_file = os.path.join(output_dir, f"{qtype}_{split}.jsonl")
        print(f"[{tag}: {qtype}] writing {len(samples)} items to {out_file}")
        self._write(samples, out_file)
        return samples

    def build_train_dataset(self, parquet_path, output_dir, qtype, desired_count, overlap_min, overlap_max, interval):
        return self._build(parquet_path, output_dir, qtype, desired_count, overlap_min, overlap_max, interval, "train")

    def build_val_dataset(self, parquet_path, output_dir, qtype, desired_count, overlap_min, overlap_max, interval):
        return self._build(parquet_path, output_dir, qtype, desired_count, overlap_min, overlap_max, interval, "val")


# ------------------------------------------------------------------------------------------------
# visual correspondence
# ------------------------------------------------------------------------------------------------
class _VisualCorrespondence(_PairTableEngine):
    stem = ""

    def _request(self, ctx, r):
        """VC-coor:277-315: swap, then the drawn members of np.intersect1d(points1, points2) as positions in it."""
        rnd, k = self.rnd, ctx["max_points_per_pair"]
        swap = rnd.random() < 0.5
        n = int(ctx["inter"][r])
        if n == 0:
            if not ctx["dry"]:
                i1, i2 = (ctx["i2"][r], ctx["i1"][r]) if swap else (ctx["i1"][r], ctx["i2"][r])
                self._warn(ctx, f"[build_training_sample] Warning: No common visible points for scene {ctx['scene'][r]} {i1}, {i2}\n")
            return None
        # random.sample(list(common), k) / random.choice(common.tolist()) draw positions; only len(common) matters
        ks = rnd.sample(range(n), k) if n >= k else [rnd.choice(range(n)) for _ in range(k)]
        return (swap, ks)

    def _names(self, ctx, r, swap):
        return (ctx["scene"][r],) + ((ctx["i2"][r], ctx["i1"][r]) if swap else (ctx["i1"][r], ctx["i2"][r]))

    def _warn_invisible(self, ctx, rec, scene_id, image1, image2):
        for bit, image in ((1, image1), (2, image2)):
            if not rec["flags"] & bit:
                self._warn(ctx, f"Warning: Point {int(rec['point_id'])} is not visible in image {image} in scene {scene_id}.\n")

    def _dataset(self, parquet_path, output_dir, desired_count, overlap_min, overlap_max, interval, visibility_info_path,
                 warning_file, max_points_per_pair, split, **extra):
        """build_train_dataset / build_val_dataset (VC-coor:401-468).  ``visibility_info_path`` is accepted for
        signature compatibility: the table is K1's bit masks, bit-identical to the file (tests/test_gpu_parity.py)."""
        tag = "Train" if split == "train" else "Val"
        df = pd.read_parquet(parquet_path)
        print(f"[{tag}] Loaded DataFrame with {len(df)} rows from {parquet_path}")
        df_sampled = sample_dataframe(df, desired_count, 0, overlap_min, overlap_max, interval)
        print(f"[{tag}] Got {len(df_sampled)} sampled rows")
        samples = [s for s in self.build_samples(df_sampled, warning_file, max_points_per_pair=max_points_per_pair, **extra) if s]
        if split == "val":
            samples = [self.convert_train_sample_to_eval_sample(s) for s in samples]
        self.rnd.shuffle(samples)
        out_file = os.path.join(output_dir, f"{split}_{self.stem}.jsonl")
        print(f"[{tag}] Writing {len(samples)} items to {out_file}")
        self._write(samples, out_file)
        return samples


class VisualCorrespondenceCoorEngine(_VisualCorrespondence):
    template_key, stem = "visual_correspondence", "visual_correspondence_coor_2_coor"

    def _finish(self, ctx, r, req, recs, dry):
        rnd = self.rnd
        swap, ks = req
        scene_id, image1, image2 = self._names(ctx, r, swap)
        conversation, p1_list, p2_list, as_predicted = [], [], [], True
        for q in range(len(ks)):
            rec = None if dry else recs[q]
            if rec is not None and rec["flags"] != 3:                            # VC-coor:325-337
                self._warn_invisible(ctx, rec, scene_id, image1, image2)
                as_predicted = False
                continue
            task = rnd.choice(self.task_description)
            question = rnd.choice(self.templates["questions"])
            answer = rnd.choice(self.templates["answers"])
            if dry:
                continue
            x1, y1, x2, y2 = (int(v) for v in rec["permille"])
            question, answer = question.format(x1=x1, y1=y1, x2=x2, y2=y2), answer.format(x1=x1, y1=y1, x2=x2, y2=y2)
            if not conversation:
                conversation = [{"from": "human", "value": f"{task}\n{question}"}, {"from": "gpt", "value": answer}]
            else:
                conversation += [{"from": "human", "value": question}, {"from": "gpt", "value": answer}]
            p1_list.append((x1, y1))
            p2_list.append((x2, y2))
        if dry:
            return None, True
        if not conversation:
            self._warn(ctx, f"[build_training_sample] Warning: No conversation for scene {scene_id} {image1}, {image2}\n")
            return None, as_predicted
        H, W = self.scene_infos.get_image_size(scene_id)
        return {"id": f"{scene_id}_{image1}_{image2}_{r}", "image": [f"{scene_id}/{image1}.jpg", f"{scene_id}/{image2}.jpg"],
                "conversations": conversation, "height_list": [H, H], "width_list": [W, W],
                "question_type": "visual_correspondence_coor_2_coor", "p1_list": p1_list, "p2_list": p2_list,
                "gt_value": list(p2_list[0])}, as_predicted

    def build_train_dataset(self, parquet_path, output_dir, desired_count, overlap_min, overlap_max, interval,
                            visibility_info_path=None, warning_file=None, max_points_per_pair=1):
        return self._dataset(parquet_path, output_dir, desired_count, overlap_min, overlap_max, interval,
                             visibility_info_path, warning_file, max_points_per_pair, "train")

    def build_val_dataset(self, parquet_path, output_dir, desired_count, overlap_min, overlap_max, interval,
                          visibility_info_path=None, warning_file=None, max_points_per_pair=1):
        assert max_points_per_pair == 1, "[Val] max_points_per_pair should be 1."
        return self._dataset(parquet_path, output_dir, desired_count, overlap_min, overlap_max, interval,
                             visibility_info_path, warning_file, max_points_per_pair, "val")


class VisualCorrespondenceDotEngine(_VisualCorrespondence):
    """VC-dot:280-428.  The two annotated JPEGs of every sample are painted on the device (nvJPEG decode,
    msp_draw_marks, nvJPEG encode) after the row loop, scene by scene."""
    template_key, stem = "visual_correspondence_dot", "visual_correspondence_dot_2_multichoice"

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.pending_marks = []                              # (scene_id, image_id, relative output path, [marks])

    def _finish(self, ctx, r, req, recs, dry):
        from .engines import _distinct_colors
        rnd = self.rnd
        swap, ks = req
        assert len(ks) == 1, "[build_training_sample] max_points_per_pair should be 1."
        scene_id, image1, image2 = self._names(ctx, r, swap)
        rec = None if dry else recs[0]
        if rec is not None and rec["flags"] != 3:                                # VC-dot:344-356
            self._warn_invisible(ctx, rec, scene_id, image1, image2)
            return None, False
        color1 = (rnd.randint(0, 255), rnd.randint(0, 255), rnd.randint(0, 255))
        correct = None if dry else (int(rec["pix"][2]), int(rec["pix"][3]))
        W, H = self.scene_infos.image_width, self.scene_infos.image_height
        incorrect, as_predicted = [], True
        while len(incorrect) < 3:                                                # VC-dot:368-373
            x, y = rnd.randint(0, W - 10), rnd.randint(0, H - 10)
            if (x, y) != correct:
                incorrect.append((x, y))
            else:
                as_predicted = False                                             # the dry run could not know this redraw
        all_points = [correct] + incorrect
        rnd.shuffle(all_points)
        labels = ["A", "B", "C", "D"]
        rnd.shuffle(labels)
        colors = _distinct_colors(rnd, 4)
        task = rnd.choice(self.task_description)
        question = rnd.choice(self.templates["questions"])
        answer = rnd.choice(self.templates["answers"])
        if dry:
            return None, True
        from .annotate import default_atlas
        atlas = default_atlas()
        labeled = dict(zip(labels, all_points))
        correct_label = [label for label, point in labeled.items() if point == correct][0]
        pt, idx = int(rec["point_id"]), r
        path1 = os.path.join(scene_id, f"{idx}_point{pt}_{image1}_{image2}_img1.jpg")
        path2 = os.path.join(scene_id, f"{idx}_point{pt}_{image1}_{image2}_img2.jpg")
        u1, v1 = int(rec["pix"][0]), int(rec["pix"][1])
        self.pending_marks.append((scene_id, image1, path1, [(u1, v1, 10, -1, 0, 0, color1, 0)]))
        self.pending_marks.append((scene_id, image2, path2, [(x, y, 10, atlas.index(label), x + 15, y, colors[i], 0)
                                                             for i, (label, (x, y)) in enumerate(labeled.items())]))
        return {"id": f"{idx}_p{pt}", "image": [path1, path2],
                "conversations": [{"from": "human", "value": f"{task}\n{question}"},
                                  {"from": "gpt", "value": answer.format(correct_label=correct_label)}],
                "height_list": [H] * 2, "width_list": [W] * 2, "question_type": "visual_correspondence_multiple_choice",
                "gt_value": correct_label, "p1_list": [u1, v1], "p2_list": [correct] + incorrect}, as_predicted

    def flush_images(self, image_output_dir, workers=4):
        """cv2.imread / circle / putText / imwrite of VC-dot:358-395 for every queued sample (annotate.write_annotated:
        every source frame decoded once, outputs painted and encoded on the device)."""
        from . import annotate
        jobs, self.pending_marks = self.pending_marks, []
        return annotate.write_annotated([((s, i), path, marks) for s, i, path, marks in jobs],
                                        lambda key: open(self.scene_infos.get_image_path(*key), "rb").read(),
                                        image_output_dir, self.device, workers)

    def _dataset_with_images(self, output_dir, split, *a, **k):
        image_output_dir = os.path.join(output_dir, "images")
        os.makedirs(image_output_dir, exist_ok=True)
        samples = self._dataset(*a, split=split, **k)
        self.flush_images(image_output_dir)
        return samples

    def build_train_dataset(self, parquet_path, output_dir, desired_count, overlap_min, overlap_max, interval,
                            visibility_info_path=None, warning_file=None, max_points_per_pair=1):
        return self._dataset_with_images(output_dir, "train", parquet_path, output_dir, desired_count, overlap_min,
                                         overlap_max, interval, visibility_info_path, warning_file, max_points_per_pair)

    def build_val_dataset(self, parquet_path, output_dir, desired_count, overlap_min, overlap_max, interval,
                          visibility_info_path=None, warning_file=None, max_points_per_pair=1):
        assert max_points_per_pair == 1, "[Val] max_points_per_pair should be 1."
        return self._dataset_with_images(output_dir, "val", parquet_path, output_dir, desired_count, overlap_min,
                                         overlap_max, interval, visibility_info_path, warning_file, max_points_per_pair)


# ------------------------------------------------------------------------------------------------
# the three main() of the reference (hard-coded paths upstream; flags here, same defaults)
# ------------------------------------------------------------------------------------------------
CAMERA_TRAIN_COUNTS = {"x_movement": 1000000, "y_movement": 1000000, "z_movement": 1000000, "yaw_movement": 1000000,
                       "pitch_movement": 1000000, "yaw_angle": 1000000, "pitch_angle": 1000000, "total_distance": 3000000,
                       "displacement_vector": 3000000}                            # CME:369-379 (this order)
CAMERA_VAL_ORDER = ("x_movement", "y_movement", "z_movement", "yaw_movement", "pitch_movement", "total_distance",
                    "yaw_angle", "pitch_angle", "displacement_vector")


def main(argv=None):
    """``python -m multi_spatialmllm_b200.pair_engines <engine>`` = the ``main()`` of CME:357-443, VC-coor:473-537,
    VC-dot:510-574: val split first, then train, seeds 0 / 1 / 2 as the reference modules set them at import."""
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("engine", choices=["camera_movement", "visual_correspondence_coor_2_coor", "visual_correspondence_dot_2_multichoice"])
    ap.add_argument("--info_path", default="data/scannet/scannet_instance_data/scenes_train_val_info_i_D5.pkl")
    ap.add_argument("--train_parquet_path", default="training_data/camera_movement/train_camera_info_D5.parquet")
    ap.add_argument("--val_parquet_path", default="evaluation_data/camera_movement/val_camera_info_D5.parquet")
    ap.add_argument("--train_output_root", default="training_data")
    ap.add_argument("--val_output_root", default="evaluation_data")
    ap.add_argument("--version", default="v1_0")
    ap.add_argument("--train_max_samples", type=int, default=None, help="per question type for camera_movement")
    ap.add_argument("--val_max_samples", type=int, default=300)
    ap.add_argument("--overlap_min", type=float, default=6)
    ap.add_argument("--overlap_max", type=float, default=35)
    ap.add_argument("--interval", type=float, default=1)
    ap.add_argument("--reference_root", default=None, help="reference checkout to read the phrasings from")
    ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args(argv)
    key = {"camera_movement": "camera_movement", "visual_correspondence_coor_2_coor": "visual_correspondence",
           "visual_correspondence_dot_2_multichoice": "visual_correspondence_dot"}[a.engine]
    seed = a.seed if a.seed is not None else {"camera_movement": 0, "visual_correspondence": 1, "visual_correspondence_dot": 2}[key]
    random.seed(seed)
    np.random.seed(seed)
    tp = load_reference_templates(a.reference_root, key) if a.reference_root else None
    train_dir = os.path.join(a.train_output_root, a.engine, a.version)
    val_dir = os.path.join(a.val_output_root, a.engine, a.version)
    os.makedirs(train_dir, exist_ok=True)
    os.makedirs(val_dir, exist_ok=True)
    bins = (a.overlap_min, a.overlap_max, a.interval)
    bins = tuple(int(v) if float(v).is_integer() else v for v in bins)
    scene_infos = SceneInfoHandler(a.info_path)
    if key == "camera_movement":
        eng = CameraMovementEngine(scene_infos, tp)
        for qtype, n_train in CAMERA_TRAIN_COUNTS.items():
            print(f"\n=== Processing question type: {qtype} ===")
            eng.build_val_dataset(a.val_parquet_path, val_dir, qtype, a.val_max_samples, *bins)
            eng.build_train_dataset(a.train_parquet_path, train_dir, qtype, a.train_max_samples or n_train, *bins)
        print("All question types processed. Done.")
        return
    cls, n_train = ((VisualCorrespondenceCoorEngine, 1000000) if key == "visual_correspondence"
                    else (VisualCorrespondenceDotEngine, 500000))
    eng = cls(scene_infos, tp)
    eng.build_val_dataset(a.val_parquet_path, val_dir, a.val_max_samples, *bins,
                          warning_file=os.path.join(val_dir, "val_warning.txt"), max_points_per_pair=1)
    eng.build_train_dataset(a.train_parquet_path, train_dir, a.train_max_samples or n_train, *bins,
                            warning_file=os.path.join(train_dir, "train_warning.txt"), max_points_per_pair=1)


if __name__ == "__main__":
    main()
