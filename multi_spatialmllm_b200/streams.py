"""Batch sample writers for the MultiSPA streams, on top of the device records.

The reference builds every sample with per-sample handler calls (two JPEG decodes, one PNG decode and one
``np.load`` of the whole scene per point, IH:284-305).  Here the host only samples ids; all geometry of a batch
is one or two kernel launches (msp_annotate_points / msp_pair_annotate) and the host formats the returned
integers into the reference's JSONL schemas:

  depth_estimation_{coor,dot}  DE-coor:242-251 / DE-dot:253-262   id, image, conversations, height_list, width_list,
                                                                   question_type, gt_value (mm), ori_coordinates
  depth_comparison_{coor,dot}  DC-coor:333-343                     + points_info[{x,y,depth,coords,letter}], is_closer_question
  visual_correspondence_coor_2_coor  VC-coor:382-393               p1_list, p2_list, gt_value
  camera_movement (9 question types)  CME:205-243                  answer_values{15 keys}, gt_value

Integer semantics are the reference's (round-half-even per-mille / mm, truncating ``int()`` pixel and mm values,
``<=`` for "closer"), verified against the reference engines' own outputs in tests/test_streams.py.  The text
templates are NOT shipped here (the reference is CC BY-NC): ``Templates.from_reference(root)`` loads them from a
user-supplied checkout; the built-in fall-back is a single neutral phrasing per stream.  Random choices use a
``numpy.random.Generator`` — the streams are statistically like the reference's, not RNG-identical; for
RNG-identical output run the reference's engine scripts with the drop-in handler.
"""
from __future__ import annotations

import importlib.util
import json
import os

import numpy as np

from .engine import DeviceScene, DeviceVideo

CAMERA_QUESTION_TYPES = ("x_movement", "y_movement", "z_movement", "yaw_movement", "pitch_movement",
                         "total_distance", "yaw_angle", "pitch_angle", "displacement_vector")


class Templates:
    """question / answer / task-description phrasings per stream."""

    NEUTRAL = {
        "depth_estimation": {"task": ["<image>\nPoint coordinates [ x , y ] are per-mille of the image width and height."],
                             "questions": ["What is the depth in mm of the point [ {x1} , {y1} ]?"],
                             "answers": ["`{depth}`"]},
        "depth_comparison": {"task": ["<image>\nPoint coordinates [ x , y ] are per-mille of the image width and height."],
                             "closer_questions": ["Which of [ {x1} , {y1} ] and [ {x2} , {y2} ] is closer to the camera?"],
                             "farther_questions": ["Which of [ {x1} , {y1} ] and [ {x2} , {y2} ] is farther from the camera?"],
                             "closer_answers": ["[ {correct_x} , {correct_y} ]"], "farther_answers": ["[ {correct_x} , {correct_y} ]"]},
        "visual_correspondence": {"task": ["Image-1: <image>\nImage-2: <image>\nCoordinates [ x , y ] are per-mille of width and height."],
                                  "questions": ["Where is the point [ {x1} , {y1} ] of Image-1 in Image-2?"],
                                  "answers": ["[ {x2} , {y2} ]"]},
        "depth_estimation_dot": {"task": ["<image>\nA point is marked with a dot."],
                                 "questions": ["What is the depth in mm of the marked point?"], "answers": ["`{depth}`"]},
        "depth_comparison_dot": {"task": ["<image>\nTwo points are marked A and B."],
                                 "closer_questions": ["Which marked point is closer to the camera?"],
                                 "farther_questions": ["Which marked point is farther from the camera?"],
                                 "closer_answers": ["{correct_label}"], "farther_answers": ["{correct_label}"]},
        "visual_correspondence_dot": {"task": ["Image-1: <image>\nImage-2: <image>\nA point is marked in Image-1; candidates are lettered in Image-2."],
                                      "questions": ["Which candidate is the marked point?"], "answers": ["{correct_label}"]},
        "camera_movement": {"task": ["Image-1: <image>\nImage-2: <image>\nDescribe the camera motion from Image-1 to Image-2."],
                            "questions": {q: [f"Report {q} of the second camera relative to the first."] for q in CAMERA_QUESTION_TYPES},
                            "answers": {q: ["`{" + q + "}`"] for q in CAMERA_QUESTION_TYPES}},
    }

    def __init__(self, tables=None):
        self.t = tables or self.NEUTRAL

    @classmethod
    def from_reference(cls, reference_root: str) -> "Templates":
        """Import the phrasings from a reference checkout (they stay in the user's copy; nothing is vendored)."""
        def load(rel, name):
            path = os.path.join(reference_root, rel)
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
        cm = load("spatial_engine/camera_movement/TEMPLATES.py", "_msp_ref_cm_templates")
        t = dict(cls.NEUTRAL)
        t["camera_movement"] = {"task": cm.TASK_DESCRIPTION, "questions": cm.QUESTION_TEMPLATES, "answers": cm.ANSWER_TEMPLATES}
        om = cls._literals(os.path.join(reference_root, "spatial_engine/object_movement/single_object_movement_engine_coord.py"),
                           ("TASK_DESCRIPTION", "QUESTION_TEMPLATES", "ANSWER_TEMPLATES"))
        for qt in om["QUESTION_TEMPLATES"]:
            t[f"object_movement:{qt}"] = {"task": om["TASK_DESCRIPTION"], "questions": om["QUESTION_TEMPLATES"][qt],
                                          "answers": om["ANSWER_TEMPLATES"][qt]}
        return cls(t)

    @staticmethod
    def _literals(path: str, names):
        """Module-level template tables of a reference file WITHOUT importing it (the engine modules seed the global RNGs
        and import optional packages when executed): only the assignments to ``names`` — plain ``NAME = literal`` and the
        follow-up ``NAME[key] = [... for q in NAME[key]]`` rewrites — are evaluated, with no builtins in scope."""
        import ast
        tree = ast.parse(open(path).read())
        ns = {"__builtins__": {}}
        for node in tree.body:
            if not (isinstance(node, ast.Assign) and len(node.targets) == 1):
                continue
            tgt = node.targets[0]
            base = tgt.value if isinstance(tgt, ast.Subscript) else tgt
            if isinstance(base, ast.Name) and base.id in names:
                exec(compile(ast.Module([node], []), path, "exec"), ns)
        missing = set(names) - set(ns)
        if missing:
            raise KeyError(f"{path}: no assignment for {sorted(missing)}")
        return {k: ns[k] for k in names}


def _pick(rng, seq):
    """One element of ``seq``: a NumPy Generator draws an index; Python's ``random`` module (or a ``random.Random``)
    is used through ``choice`` so that the draw consumes the reference's RNG stream exactly (it calls random.choice)."""
    if hasattr(rng, "integers"):
        return seq[int(rng.integers(0, len(seq)))]
    return rng.choice(seq)


def _conversation(task, question, answer):
    return [{"from": "human", "value": f"{task}\n{question}"}, {"from": "gpt", "value": answer}]


# ------------------------------------------------------------------------------------------------
# depth perception
# ------------------------------------------------------------------------------------------------
def annotate(sc: DeviceScene, frames, points):
    a = sc.annotate_points(np.asarray(points, np.int32), np.asarray(frames, np.int32))
    return {k: v.cpu().numpy() for k, v in a.items()}


def depth_estimation_samples(sc, scene_id, image_ids, frames, points, variant="coor", templates=None, rng=None):
    """One sample per (frame, point); invisible points are skipped like DE-coor:207-213."""
    tp, rng = (templates or Templates()).t["depth_estimation"], rng or np.random.default_rng(0)
    H, W = sc.image_hw
    a = annotate(sc, frames, points)
    out = []
    for i, (f, p) in enumerate(zip(frames, points)):
        if not a["visible"][i]:
            continue
        x, y = (int(v) for v in a["permille"][i])
        depth, (u, v) = int(a["depth_mm"][i]), (int(c) for c in a["pix"][i])
        image_id = image_ids[f]
        q = _pick(rng, tp["questions"]).format(x1=x, y1=y)
        ans = _pick(rng, tp["answers"]).format(x1=x, y1=y, depth=depth)
        img = f"{scene_id}/{image_id}.jpg" if variant == "coor" else f"{scene_id}/{image_id}_p{int(p)}_annotated.jpg"
        out.append({"id": f"{scene_id}_{image_id}_point{int(p)}", "image": [img],
                    "conversations": _conversation(_pick(rng, tp["task"]), q, ans),
                    "height_list": [H], "width_list": [W], "question_type": f"depth_estimation_{variant}",
                    "gt_value": depth, "ori_coordinates": [u, v]})
    return out


def depth_comparison_samples(sc, scene_id, image_ids, frames, point_pairs, closer_question, variant="coor", templates=None, rng=None):
    """One sample per (frame, (pA, pB)); equal mm depths are rejected (DC-coor:289-294), "closer" uses <= (DC-coor:304)."""
    tp, rng = (templates or Templates()).t["depth_comparison"], rng or np.random.default_rng(0)
    H, W = sc.image_hw
    pp = np.asarray(point_pairs, np.int32).reshape(-1, 2)
    fr2 = np.repeat(np.asarray(frames, np.int32), 2)
    a = annotate(sc, fr2, pp.reshape(-1))
    out = []
    for i, f in enumerate(frames):
        info = []
        for s in range(2):
            k = 2 * i + s
            if a["visible"][k]:
                info.append({"x": int(a["permille"][k][0]), "y": int(a["permille"][k][1]), "depth": int(a["depth_mm"][k]),
                             "coords": (int(a["pix"][k][0]), int(a["pix"][k][1])), "letter": "AB"[s]})
        if len(info) != 2 or info[0]["depth"] == info[1]["depth"]:
            continue
        p1, p2 = info
        closer, farther = (p1, p2) if p1["depth"] <= p2["depth"] else (p2, p1)
        is_closer = bool(closer_question[i])
        pick = closer if is_closer else farther
        kind = "closer" if is_closer else "farther"
        q = _pick(rng, tp[f"{kind}_questions"]).format(x1=p1["x"], y1=p1["y"], x2=p2["x"], y2=p2["y"])
        ans = _pick(rng, tp[f"{kind}_answers"]).format(correct_x=pick["x"], correct_y=pick["y"])
        image_id = image_ids[f]
        out.append({"id": f"{scene_id}_{image_id}_p{int(pp[i][0])}_p{int(pp[i][1])}", "image": [f"{scene_id}/{image_id}.jpg"],
                    "conversations": _conversation(_pick(rng, tp["task"]), q, ans), "height_list": [H], "width_list": [W],
                    "question_type": "depth_comparison_coordinate" if variant == "coor" else "depth_comparison_dot",
                    "gt_value": [pick["x"], pick["y"]], "points_info": info, "is_closer_question": is_closer})
    return out


# ------------------------------------------------------------------------------------------------
# pair streams
# ------------------------------------------------------------------------------------------------
def visual_correspondence_samples(sc, scene_id, image_ids, pairs, sel, sel_mode=1, start_idx=0, templates=None, rng=None):
    """One sample per ordered pair (image1, image2); ``sel`` picks the common visible point (VC-coor:303-315)."""
    tp, rng = (templates or Templates()).t["visual_correspondence"], rng or np.random.default_rng(0)
    H, W = sc.image_hw
    rec = sc.records_to_numpy(sc.pair_annotate(np.asarray(pairs, np.int32), sel, sel_mode))
    out = []
    for i, ((a, b), r) in enumerate(zip(np.asarray(pairs).tolist(), rec)):
        if r["point_id"] < 0 or r["flags"] != 3:          # no common point / failed re-check (VC-coor:304-337)
            continue
        x1, y1, x2, y2 = (int(v) for v in r["permille"])
        i1, i2 = image_ids[a], image_ids[b]
        q = _pick(rng, tp["questions"]).format(x1=x1, y1=y1, x2=x2, y2=y2)
        ans = _pick(rng, tp["answers"]).format(x1=x1, y1=y1, x2=x2, y2=y2)
        out.append({"id": f"{scene_id}_{i1}_{i2}_{start_idx + i}", "image": [f"{scene_id}/{i1}.jpg", f"{scene_id}/{i2}.jpg"],
                    "conversations": _conversation(_pick(rng, tp["task"]), q, ans), "height_list": [H, H], "width_list": [W, W],
                    "question_type": "visual_correspondence_coor_2_coor", "p1_list": [(x1, y1)], "p2_list": [(x2, y2)],
                    "gt_value": [x2, y2]})
    return out


# ------------------------------------------------------------------------------------------------
# `*_dot` variants: same geometry, plus the marks `annotate.draw_marks` paints on the colour frames
# ------------------------------------------------------------------------------------------------
DOT_RADIUS = 10                                            # DE-dot:171, DC-dot:337, VC-dot:363,390


def _mark(marks, image, xy, color, letter=None, org=None, atlas=None, radius=DOT_RADIUS):
    marks.append((image, int(xy[0]), int(xy[1]), radius, -1 if letter is None else atlas.index(letter),
                  0 if org is None else int(org[0]), 0 if org is None else int(org[1]), tuple(int(c) for c in color), 0))


def _finish_marks(marks):
    from .annotate import MARK_DTYPE
    return np.array(marks, dtype=MARK_DTYPE) if marks else np.zeros(0, MARK_DTYPE)


def depth_estimation_dot_samples(sc, scene_id, image_ids, frames, points, templates=None, rng=None):
    """DE-dot:197-262.  Returns (samples, marks, plan): output image k of the batch is a copy of colour frame
    ``plan[k]`` with ``marks[marks["image"] == k]`` painted on it and is stored under ``samples[k]["image"][0]``."""
    from .annotate import distinct_colors
    tp, rng = (templates or Templates()).t["depth_estimation_dot"], rng or np.random.default_rng(0)
    H, W = sc.image_hw
    a = annotate(sc, frames, points)
    out, marks, plan = [], [], []
    for i, (f, p) in enumerate(zip(frames, points)):
        if not a["visible"][i]:
            continue
        x, y = (int(v) for v in a["permille"][i])
        depth, (u, v) = int(a["depth_mm"][i]), (int(c) for c in a["pix"][i])
        image_id = image_ids[f]
        _mark(marks, len(plan), (u, v), distinct_colors(1, rng)[0])
        plan.append(int(f))
        ans = _pick(rng, tp["answers"]).format(x1=x, y1=y, depth=depth)
        out.append({"id": f"{scene_id}_{image_id}_point{int(p)}", "image": [f"{scene_id}/{image_id}_p{int(p)}_annotated.jpg"],
                    "conversations": _conversation(_pick(rng, tp["task"]), _pick(rng, tp["questions"]), ans),
                    "height_list": [H], "width_list": [W], "question_type": "depth_estimation_dot",
                    "gt_value": depth, "ori_coordinates": [u, v]})
    return out, _finish_marks(marks), plan


def depth_comparison_dot_samples(sc, scene_id, image_ids, frames, point_pairs, closer_question, templates=None, rng=None,
                                 atlas=None):
    """DC-dot:262-373: letters A/B dealt at random, a random colour per point, label at (x+15, y+15), gt = letter."""
    from .annotate import default_atlas
    tp, rng = (templates or Templates()).t["depth_comparison_dot"], rng or np.random.default_rng(0)
    atlas = atlas or default_atlas()
    H, W = sc.image_hw
    pp = np.asarray(point_pairs, np.int32).reshape(-1, 2)
    a = annotate(sc, np.repeat(np.asarray(frames, np.int32), 2), pp.reshape(-1))
    out, marks, plan = [], [], []
    for i, f in enumerate(frames):
        info = []
        for s in range(2):
            k = 2 * i + s
            if a["visible"][k]:
                info.append({"x": int(a["permille"][k][0]), "y": int(a["permille"][k][1]), "depth": int(a["depth_mm"][k]),
                             "coords": (int(a["pix"][k][0]), int(a["pix"][k][1]))})
        if len(info) != 2 or info[0]["depth"] == info[1]["depth"]:
            continue
        letters = ["A", "B"] if rng.random() < 0.5 else ["B", "A"]
        if rng.random() < 0.5:
            info = info[::-1]
        for pi, letter in zip(info, letters):
            pi["letter"] = letter
        p1, p2 = info
        closer, farther = (p1, p2) if p1["depth"] <= p2["depth"] else (p2, p1)
        is_closer = bool(closer_question[i])
        pick = closer if is_closer else farther
        kind = "closer" if is_closer else "farther"
        for pi in info:
            cx, cy = pi["coords"]
            _mark(marks, len(plan), (cx, cy), rng.integers(0, 256, 3), pi["letter"], (cx + 15, cy + 15), atlas)
        plan.append(int(f))
        image_id = image_ids[f]
        stem = f"{image_id}_p{int(pp[i][0])}_p{int(pp[i][1])}"
        out.append({"id": f"{scene_id}_{stem}", "image": [f"{scene_id}/{stem}_annotated.jpg"],
                    "conversations": _conversation(_pick(rng, tp["task"]), _pick(rng, tp[f"{kind}_questions"]),
                                                   _pick(rng, tp[f"{kind}_answers"]).format(correct_label=pick["letter"])),
                    "height_list": [H], "width_list": [W], "question_type": "depth_comparison_annotated",
                    "gt_value": pick["letter"], "points_info": info, "is_closer_question": is_closer})
    return out, _finish_marks(marks), plan


def visual_correspondence_dot_samples(sc, scene_id, image_ids, pairs, sel, sel_mode=1, start_idx=0, templates=None, rng=None,
                                      atlas=None):
    """VC-dot:357-426: the point dotted in image 1; in image 2 the true location and three random decoys
    (x in [0, W-10], y in [0, H-10], != the true pixel), shuffled, lettered A-D at random, distinct colours,
    labels at (x+15, y).  Two output images per sample (plan holds both frames)."""
    from .annotate import default_atlas, distinct_colors
    tp, rng = (templates or Templates()).t["visual_correspondence_dot"], rng or np.random.default_rng(0)
    atlas = atlas or default_atlas()
    H, W = sc.image_hw
    rec = sc.records_to_numpy(sc.pair_annotate(np.asarray(pairs, np.int32), sel, sel_mode))
    out, marks, plan = [], [], []
    for i, ((fa, fb), r) in enumerate(zip(np.asarray(pairs).tolist(), rec)):
        if r["point_id"] < 0 or r["flags"] != 3:
            continue
        u1, v1, u2, v2 = (int(c) for c in r["pix"])
        correct = (u2, v2)
        decoys = []
        while len(decoys) < 3:
            cand = (int(rng.integers(0, W - 10 + 1)), int(rng.integers(0, H - 10 + 1)))
            if cand != correct:
                decoys.append(cand)
        pts = [correct] + decoys
        order = rng.permutation(4)
        labels = [("A", "B", "C", "D")[k] for k in rng.permutation(4)]
        placed = [(labels[k], pts[order[k]]) for k in range(4)]
        correct_label = next(lab for lab, pt in placed if pt == correct)
        colors = distinct_colors(4, rng)
        _mark(marks, len(plan), (u1, v1), rng.integers(0, 256, 3))
        plan.append(int(fa))
        for (lab, (x, y)), col in zip(placed, colors):
            _mark(marks, len(plan), (x, y), col, lab, (x + 15, y), atlas)
        plan.append(int(fb))
        idx, pt, i1, i2 = start_idx + i, int(r["point_id"]), image_ids[fa], image_ids[fb]
        out.append({"id": f"{idx}_p{pt}",
                    "image": [f"{scene_id}/{idx}_point{pt}_{i1}_{i2}_img1.jpg", f"{scene_id}/{idx}_point{pt}_{i1}_{i2}_img2.jpg"],
                    "conversations": _conversation(_pick(rng, tp["task"]), _pick(rng, tp["questions"]),
                                                   _pick(rng, tp["answers"]).format(correct_label=correct_label)),
                    "height_list": [H, H], "width_list": [W, W], "question_type": "visual_correspondence_multiple_choice",
                    "gt_value": correct_label, "p1_list": [u1, v1], "p2_list": [correct] + decoys})
    return out, _finish_marks(marks), plan


def render_annotated(frames_bgr, plan, marks, device="cuda:0", atlas=None):
    """Gather ``frames_bgr[plan]`` ((F,H,W,3) uint8, host or device) into a device batch and paint the marks."""
    import torch
    from .annotate import draw_marks
    src = frames_bgr if isinstance(frames_bgr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames_bgr))
    idx = torch.as_tensor(np.asarray(plan, np.int64))
    batch = src.to(device)[idx.to(device)].contiguous() if src.is_cuda or src.numel() < (1 << 28) else src[idx].to(device)
    return draw_marks(batch, marks, atlas)


def camera_answer_values(rel_disp, yaw_angle, pitch_angle):
    """CME:168-172 (yaw wrap) and CME:205-225 (signs, truncating ints) from the device record of the ORDERED pair."""
    if abs(yaw_angle) > 180:
        yaw_angle = yaw_angle - 360 if yaw_angle > 0 else yaw_angle + 360
    d = np.asarray(rel_disp, dtype=np.float64)
    return {
        "x_movement": "right" if d[0] > 0 else "left", "y_movement": "down" if d[1] > 0 else "up",
        "z_movement": "forward" if d[2] > 0 else "backward",
        "yaw_movement": "left" if yaw_angle > 0 else "right", "pitch_movement": "up" if pitch_angle > 0 else "down",
        "x_distance": int(abs(d[0]) * 1000), "y_distance": int(abs(d[1]) * 1000), "z_distance": int(abs(d[2]) * 1000),
        "yaw_angle": int(abs(yaw_angle)), "pitch_angle": int(abs(pitch_angle)),
        "x_value": int(d[0] * 1000), "y_value": int(d[1] * 1000), "z_value": int(d[2] * 1000),
        "total_distance": int(np.linalg.norm(d) * 1000), "displacement_vector": d.tolist(),
    }


def camera_movement_samples(sc, scene_id, image_ids, pairs, question_type, start_idx=0, templates=None, rng=None):
    """One sample per ORDERED pair (image1 -> image2); swap a pair by passing it reversed (CME:163-167)."""
    assert question_type in CAMERA_QUESTION_TYPES
    tp, rng = (templates or Templates()).t["camera_movement"], rng or np.random.default_rng(0)
    H, W = sc.image_hw
    pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
    rec = sc.records_to_numpy(sc.pair_annotate(pairs, np.zeros(len(pairs), np.uint32), 0))
    out = []
    for i, ((a, b), r) in enumerate(zip(pairs.tolist(), rec)):
        if not (r["overlap"] >= 0.1):                     # CME:199-201: low / undefined overlap is not supported
            continue
        av = camera_answer_values(r["rel_disp"], float(r["dyaw"]), float(r["dpitch"]))
        q = _pick(rng, tp["questions"][question_type])
        ans = _pick(rng, tp["answers"][question_type]).format(**av)
        out.append({"id": start_idx + i, "image": [f"{scene_id}/{image_ids[a]}.jpg", f"{scene_id}/{image_ids[b]}.jpg"],
                    "conversations": _conversation(_pick(rng, tp["task"]), q, ans), "height_list": [H, H], "width_list": [W, W],
                    "answer_values": av, "question_type": question_type, "gt_value": av[question_type]})
    return out


def _randint(rng, lo, hi):
    """random.randint(lo, hi) on either kind of generator (inclusive bounds)."""
    return int(rng.integers(lo, hi + 1)) if hasattr(rng, "integers") else rng.randint(lo, hi)


def _object_movement(vid: DeviceVideo, video_id, triples, question_type, object_thr, camera_thr, templates, rng, dot, seen):
    tabs = (templates or Templates()).t
    tp = tabs.get(f"object_movement:{question_type}") or tabs.get("object_movement", {
        "task": ["Image-1: <image>\nImage-2: <image>\nCoordinates [ x , y ] are per-mille of width and height."],
        "questions": ["How far did the point at [ {x1} , {y1} ] in Image-1 move (mm), in camera-1 axes?"],
        "answers": ["`{total_distance}` [ {x_value} , {y_value} , {z_value} ]"]})
    rng = rng or np.random.default_rng(0)
    H, W = vid.image_hw
    tri = np.asarray(triples, np.int32).reshape(-1, 3)
    disp = vid.pair_displacement(tri).cpu().numpy()
    pm, ok = vid.permille.cpu().numpy(), vid.valid.cpu().numpy().astype(bool)
    uvn = vid.uvn.cpu().numpy() if dot else None
    w2c_h = vid.w2c.cpu().numpy().reshape(-1, 4, 4)
    out, marks, plan = [], [], []
    for (f1, f2, pt), d in zip(tri.tolist(), disp):
        if not (ok[f1, pt] and ok[f2, pt]):
            continue
        moving = bool(d[0] >= object_thr)
        # a static point has its displacement zeroed BEFORE the rotation into camera-1 axes (OM-coord:337-341, 354-356):
        # the product with a zero vector keeps IEEE signed zeros, which json.dumps shows ("-0.0")
        dist, dcam = (float(d[0]), d[1:4].copy()) if moving else (0, (w2c_h[f1] @ np.zeros(4))[:3])
        x1, y1 = (int(v) for v in pm[f1, pt])
        x2, y2 = (int(v) for v in pm[f2, pt])
        task = _pick(rng, tp["task"])                      # draw order of OM-coord:367-376: task, question, answer
        q = _pick(rng, tp["questions"]).format(x1=x1, y1=y1)
        ans = _pick(rng, tp["answers"]).format(total_distance=round(dist * 1000), x_value=round(float(dcam[0]) * 1000),
                                               y_value=round(float(dcam[1]) * 1000), z_value=round(float(dcam[2]) * 1000))
        if not moving:
            ans = "The point did not move. " + ans
        if dot:
            # OM-dot:408-425: image 1 gets one filled circle of radius W // 100 at (int(u_n * W), int(v_n * H)) in a colour
            # drawn from the RNG the first time this (frame, point) is annotated; image 2 is a plain copy
            if (f1, pt) not in seen:
                seen.add((f1, pt))
                color = tuple(_randint(rng, 0, 255) for _ in range(3))                 # generate_distinct_colors(1), OM-dot:38-49
                _mark(marks, len(plan), (int(uvn[f1, pt, 0] * W), int(uvn[f1, pt, 1] * H)), color, radius=W // 100)
                plan.append((f1, f"{video_id}/{f1:05d}_{pt}_annotated.jpg"))
            images = [f"{video_id}/{f1:05d}_{pt}_annotated.jpg", f"{video_id}/{f2:05d}.jpg"]
            vector_gt = question_type != "total_distance"   # OM-dot:435 tests the bare name: always the vector for tapvid3d_* types
        else:
            images = [f"{video_id}/{f1:05d}.jpg", f"{video_id}/{f2:05d}.jpg"]
            vector_gt = "total_distance" not in question_type
        out.append({"id": f"{video_id}_{f1}_{f2}_{pt}" + ("_ann" if dot else ""), "image": images,
                    "conversations": _conversation(task, q, ans), "height_list": [H, H], "width_list": [W, W],
                    "gt_value": dcam.tolist() if vector_gt else int(dist * 1000),
                    "question_type": question_type, "point_moving": int(moving), "cam_moving": int(d[4] >= camera_thr),
                    "p1": (x1, y1), "p2": (x2, y2)})
    return out, marks, plan


def object_movement_samples(vid: DeviceVideo, video_id, triples, question_type="tapvid3d_total_distance",
                            object_thr=0.01, camera_thr=0.01, templates=None, rng=None):
    """TAPVid-3D two-frame samples for (frame1, frame2, track) triples (OM-coord:317-404).

    Static points (||dw|| < object_thr) get distance 0 and a zero vector (OM-coord:337-341); a triple whose point
    does not project inside either frame is skipped (OM-coord:360-362).  ``rng``: a NumPy Generator, or Python's
    ``random`` module / a ``random.Random`` to consume the reference's RNG stream draw for draw.
    """
    return _object_movement(vid, video_id, triples, question_type, object_thr, camera_thr, templates, rng, False, None)[0]


def object_movement_dot_samples(vid: DeviceVideo, video_id, triples, question_type="tapvid3d_total_distance",
                                object_thr=0.01, camera_thr=0.01, templates=None, rng=None, seen=None):
    """The dot variant (single_object_movement_engine_dot.py:341-441).  Returns (samples, marks, plan): annotated image k
    is frame ``plan[k][0]`` with ``marks[marks["image"] == k]`` painted on it, to be stored as ``plan[k][1]``; ``seen``
    carries the (frame, point) images already annotated (the reference skips files that exist, OM-dot:417-421)."""
    seen = set() if seen is None else seen
    out, marks, plan = _object_movement(vid, video_id, triples, question_type, object_thr, camera_thr, templates, rng, True, seen)
    return out, _finish_marks(marks), plan


def object_perception_samples(scene_id, image_hw, dimension_name, object_dims_m, object_categories, coverage_by_object,
                              max_k=6, templates=None, rng=None):
    """single_object_perception_engine.build_lwh_qa_samples (SOPE:129-227) for one scene.

    coverage_by_object: {object_id: {k: [image-id tuples]}} for this dimension (tables.object_coverage_scene);
    object_dims_m / object_categories: {object_id: metres} / {object_id: raw category}.  gt_value = int(round(m*1000)).
    Returns {k: [samples]}; the image order inside a combination is shuffled like the reference (SOPE:188)."""
    tp = (templates or Templates()).t.get("object_perception", {
        "task": ["Use the images to estimate the size of the object."],
        "questions": ["What is the {dimension} of the {object_category} in millimetres?"],
        "answers": ["`{value_mm}`"]})
    rng = rng or np.random.default_rng(0)
    H, W = image_hw
    out = {k: [] for k in range(1, max_k + 1)}
    for object_id, k_dict in coverage_by_object.items():
        val_mm = int(round(float(object_dims_m[object_id]) * 1000))
        cat = object_categories[object_id]
        for k_key, combos in k_dict.items():
            k = int(k_key)
            if k < 1 or k > max_k:
                continue
            for ci, combo in enumerate(combos):
                if not combo:
                    continue
                combo = [combo[i] for i in rng.permutation(len(combo))]
                prefix = "\n".join(f"Image-{i}: <image>" for i in range(1, len(combo) + 1))
                q = _pick(rng, tp["questions"]).format(dimension=dimension_name, object_category=cat)
                ans = _pick(rng, tp["answers"]).format(dimension=dimension_name, value_mm=val_mm, object_category=cat)
                out[k].append({"id": f"{scene_id}_{object_id}_{k}_{ci}", "image": [f"{scene_id}/{img}.jpg" for img in combo],
                               "conversations": [{"from": "human", "value": f"{prefix}\n{_pick(rng, tp['task'])}\n{q}"},
                                                 {"from": "gpt", "value": ans}],
                               "height_list": [H] * len(combo), "width_list": [W] * len(combo),
                               "question_type": f"object_perception_{dimension_name}_estimation", "gt_value": val_mm})
    return out


def to_eval_sample(sample: dict) -> dict:
    """DE-coor / VC-coor / CME convert_train_sample_to_eval_sample: the first human turn becomes ``text``."""
    s = dict(sample)
    s["text"] = s.pop("conversations")[0]["value"]
    return s


def write_jsonl(samples, path: str) -> int:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        for s in samples:
            f.write(json.dumps(s) + "\n")
    return len(samples)
