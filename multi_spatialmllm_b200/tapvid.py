"""TAPVid-3D object_movement: the per-video sample generation of the reference
(single_object_movement_engine_coord.py:406-575, ``generate_qa_training_single_scene``) on the GPU path.

  DeviceVideo.project                 world transform + pinhole projection of every (frame, track)     OM-coord:446-451, 293-315
  DeviceVideo.rigid_body_groups       O(T N^2) distance-change accumulation on the device, linkage on host  OM-coord:49-105
  sample_video_pairs                  all visible frame pairs of the chosen tracks -> distances on the device
                                      (msp_track_all_pair_distances), static / moving split, 10-bin histogram sampling,
                                      reverse augmentation                                              OM-coord:468-567
  streams.object_movement_samples     two-frame samples in the reference JSONL schema                    OM-coord:317-404

``sample_video_pairs`` consumes Python's ``random`` stream call for call like the reference (shuffle per group,
random.choice of a static pair, random.sample per over-full bin, random.sample for the augmentation), so with the same
seed it selects the same (point, frame1, frame2) triples; the Python loops of the reference (pair enumeration, per-pair
np.digitize, list sorting) are replaced by array operations on the device distances.
"""
from __future__ import annotations

import collections.abc
import os
import random

import numpy as np
import torch

from . import streams
from ._lib import check, lib
from .engine import DeviceVideo, _ptr, _stream

OBJECT_NOT_MOVING_THRESHOLD = 0.01      # OM-coord:288
CAMERA_NOT_MOVING_THRESHOLD = 0.01      # OM-coord:289


class _Pairs(collections.abc.Sequence):
    """(distance, frame1, frame2) tuples backed by arrays: random.choice / random.sample only need len() and indexing,
    so the lists the reference builds in Python never have to exist."""

    def __init__(self, dist, f1, f2):
        self.dist, self.f1, self.f2 = dist, f1, f2

    def __len__(self):
        return len(self.dist)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return _Pairs(self.dist[i], self.f1[i], self.f2[i])
        return (self.dist[i], self.f1[i], self.f2[i])


def all_pair_distances(vid: DeviceVideo, track_ids, visible_frames):
    """For each track its i<j visible-frame pairs (i-major, OM-coord:485-486) and their world-space distances
    (OM-coord:492-495) computed on the device.  Returns a list of (dist float64 (P,), frame1 (P,), frame2 (P,))."""
    if vid.world is None:
        vid.project()
    nt = len(track_ids)
    if nt == 0:
        return []
    counts = np.array([len(v) for v in visible_frames], np.int64)
    foff = np.zeros(nt + 1, np.int64)
    foff[1:] = np.cumsum(counts)
    npairs = counts * (counts - 1) // 2
    poff = np.zeros(nt + 1, np.int64)
    poff[1:] = np.cumsum(npairs)
    dev = vid.device
    dist = torch.empty(max(int(poff[-1]), 1), dtype=torch.float64, device=dev)
    if poff[-1] > 0:
        frames = torch.from_numpy(np.concatenate(visible_frames).astype(np.int32)).to(dev)
        tid = torch.from_numpy(np.asarray(track_ids, np.int32)).to(dev)
        foff_d, poff_d = torch.from_numpy(foff).to(dev), torch.from_numpy(poff).to(dev)
        check(lib.msp_track_all_pair_distances(_ptr(vid.world), vid.T, vid.N, _ptr(tid), nt, _ptr(frames), _ptr(foff_d),
                                               _ptr(poff_d), int(npairs.max()), _ptr(dist), _stream()),
              "msp_track_all_pair_distances")
    dist_h = dist.cpu().numpy()
    out = []
    for t in range(nt):
        vf = np.asarray(visible_frames[t])
        ii, jj = np.triu_indices(len(vf), k=1)
        out.append((dist_h[poff[t]:poff[t + 1]], vf[ii], vf[jj]))
    return out


def sample_video_pairs(vid: DeviceVideo, visibility, groups, npoints_per_group=5, npairs_per_bin=1e8, augment=True,
                       augment_ratio=1.0, object_thr=OBJECT_NOT_MOVING_THRESHOLD, rnd=random, future_frame_windows=None):
    """OM-coord:468-567.  ``groups`` is shuffled in place like the reference's; ``rnd`` is the ``random`` module or a
    ``random.Random``.  Returns [{"point_index", "frame1", "frame2"}, ...] in the reference's order.

    Kept quirks: ``npairs_per_bin`` is overwritten by each track's middle-bin count and carries over to the following
    tracks (OM-coord:535-538); the ``future_frame_windows`` test unpacks (dist, frame1, frame2) as ``frame1, frame2 =
    disp[0], disp[1]`` and therefore compares the FIRST FRAME INDEX with ``dist + window`` (OM-coord:505-508,
    OM-dot:550-552).  With the default window of 1e8 it never fires; the dot engine on ``drivetrack`` sets 5
    (OM-dot:303) and then drops every pair whose first frame index exceeds ``dist + 5`` — replicated when
    ``future_frame_windows`` is given."""
    visibility = np.asarray(visibility)
    sample_pairs = []
    for group in groups:
        rnd.shuffle(group)
        selected = group[:npoints_per_group]
        vis = [np.where(visibility[:, p])[0] for p in selected]
        keep = [k for k, v in enumerate(vis) if len(v) >= 2]
        dists = all_pair_distances(vid, [selected[k] for k in keep], [vis[k] for k in keep])
        for k, (d, f1, f2) in zip(keep, dists):
            point_idx = selected[k]
            if future_frame_windows is not None:
                ok = ~(f1 > d + future_frame_windows)                    # the reference's mis-unpacked window test
                d, f1, f2 = d[ok], f1[ok], f2[ok]
                if len(d) == 0:
                    continue
            static = d < object_thr
            chosen = []
            if static.any():
                idx = np.flatnonzero(static)
                chosen.append(rnd.choice(_Pairs(d[idx], f1[idx], f2[idx])))
            if not static.all():
                idx = np.flatnonzero(~static)
                order = idx[np.argsort(d[idx], kind="stable")]          # list.sort(key=dist) is stable
                ds, a, b = d[order], f1[order], f2[order]
                edges = np.histogram_bin_edges(ds, bins=10)
                # np.digitize(x, edges) - 1 capped at 9; the distances are sorted, so every bin is a contiguous run
                cuts = np.concatenate([[0], np.searchsorted(ds, edges[1:10], side="left"), [len(ds)]])
                bins = [_Pairs(ds[cuts[q]:cuts[q + 1]], a[cuts[q]:cuts[q + 1]], b[cuts[q]:cuts[q + 1]]) for q in range(10)]
                npairs_per_bin = max(min(len(bins[4]), npairs_per_bin), 1)
                for bn in bins:
                    chosen.extend(rnd.sample(bn, npairs_per_bin) if len(bn) > npairs_per_bin else list(bn))
            for _, fa, fb in chosen:
                sample_pairs.append({"point_index": point_idx, "frame1": fa, "frame2": fb})
    if augment:
        for s in rnd.sample(sample_pairs, int(len(sample_pairs) * augment_ratio)):
            sample_pairs.append({"point_index": s["point_index"], "frame1": s["frame2"], "frame2": s["frame1"]})
    return sample_pairs


def load_video(npz_path: str):
    """The npz keys of the TAPVid-3D release (reference README.md:112-123, OM-coord:441-444)."""
    gt = np.load(npz_path, allow_pickle=True)
    if "extrinsics_w2c" not in gt.files:
        raise NotImplementedError(f"{npz_path}: no extrinsics_w2c (camera-frame-only clips are not supported; ADT and "
                                  "PStudio, the two subsets the reference lists, carry extrinsics)")
    return {"images_jpeg_bytes": gt["images_jpeg_bytes"], "fx_fy_cx_cy": gt["fx_fy_cx_cy"], "tracks_XYZ": gt["tracks_XYZ"],
            "visibility": gt["visibility"], "extrinsics_w2c": gt["extrinsics_w2c"]}


def export_frames(images_jpeg_bytes, scene_img_dir: str):
    """OM-coord:417-438: frames as %05d.jpg (skipped when the directory already holds them all).  The JPEG streams are
    written verbatim (the reference decodes and re-encodes each one); returns (height, width)."""
    import cv2
    os.makedirs(scene_img_dir, exist_ok=True)
    have = [f for f in os.listdir(scene_img_dir) if f.endswith(".jpg")]
    if len(have) != len(images_jpeg_bytes):
        for i, frame_bytes in enumerate(images_jpeg_bytes):
            with open(os.path.join(scene_img_dir, f"{i:05d}.jpg"), "wb") as f:
                f.write(bytes(frame_bytes))
    img = cv2.imdecode(np.frombuffer(bytes(images_jpeg_bytes[0]), np.uint8), cv2.IMREAD_UNCHANGED)
    return int(img.shape[0]), int(img.shape[1])


def render_dot_images(video, video_id: str, samples, marks, plan, img_output_dir: str, device="cuda:0", batch=64) -> int:
    """OM-dot:408-425: write ``<frame1>_<point>_annotated.jpg`` (decode on nvJPEG, one filled circle painted by
    msp_draw_marks, encode on nvJPEG) and a plain copy of every second frame.  Returns the number of files written."""
    from . import annotate
    os.makedirs(os.path.join(img_output_dir, video_id), exist_ok=True)
    per_image = [[] for _ in plan]
    for m in marks:
        per_image[int(m["image"])].append((int(m["x"]), int(m["y"]), int(m["radius"]), int(m["glyph"]), int(m["org_x"]),
                                           int(m["org_y"]), tuple(int(c) for c in m["bgr"]), 0))
    written = annotate.write_annotated([(f, rel, per_image[k]) for k, (f, rel) in enumerate(plan)],
                                       lambda f: bytes(video["images_jpeg_bytes"][f]), img_output_dir, device)
    for s in samples:                                        # image 2 is not annotated: the original JPEG stream, verbatim
        path = os.path.join(img_output_dir, s["image"][1])
        if not os.path.exists(path):
            frame = int(os.path.basename(path).split(".")[0])
            with open(path, "wb") as f:
                f.write(bytes(video["images_jpeg_bytes"][frame]))
            written += 1
    return written


def generate_video_samples(video, video_id: str, question_type="tapvid3d_total_distance", npoints_per_group=5,
                           npairs_per_bin=1e8, img_output_dir=None, image_hw=None, augment=True, augment_ratio=1.0,
                           templates=None, rnd=random, device="cuda:0", dot=False, object_thr=OBJECT_NOT_MOVING_THRESHOLD,
                           camera_thr=CAMERA_NOT_MOVING_THRESHOLD, future_frame_windows=None):
    """generate_qa_training_single_scene for one video (a dict as returned by ``load_video``): OM-coord:406-575, or with
    ``dot=True`` single_object_movement_engine_dot.py:448-616 (annotated first image, ids ending in ``_ann``; the
    annotated JPEGs go to ``img_output_dir``).  Drivetrack's looser thresholds (OM-dot:301-302) are the two ``*_thr``."""
    if img_output_dir is not None and not dot:
        image_hw = export_frames(video["images_jpeg_bytes"], os.path.join(img_output_dir, video_id))
    elif image_hw is None:
        import cv2
        img = cv2.imdecode(np.frombuffer(bytes(video["images_jpeg_bytes"][0]), np.uint8), cv2.IMREAD_UNCHANGED)
        image_hw = (int(img.shape[0]), int(img.shape[1]))
    tracks = np.asarray(video["tracks_XYZ"])
    vid = DeviceVideo(tracks.astype(np.float32), video["extrinsics_w2c"], video["fx_fy_cx_cy"], image_hw, device=device)
    vid.project()
    groups, _ = vid.rigid_body_groups()                               # OM-coord:454-455 (threshold 0.1, groups > 5 points)
    pairs = sample_video_pairs(vid, video["visibility"], groups, npoints_per_group, npairs_per_bin, augment, augment_ratio,
                               object_thr=object_thr, rnd=rnd, future_frame_windows=future_frame_windows)
    triples = [(int(p["frame1"]), int(p["frame2"]), int(p["point_index"])) for p in pairs]
    if not dot:
        return streams.object_movement_samples(vid, video_id, triples, question_type, object_thr, camera_thr,
                                               templates=templates, rng=rnd)
    samples, marks, plan = streams.object_movement_dot_samples(vid, video_id, triples, question_type, object_thr, camera_thr,
                                                               templates=templates, rng=rnd)
    if img_output_dir is not None:
        render_dot_images(video, video_id, samples, marks, plan, img_output_dir, device)
    return samples


def generate_split(npz_files, output_file: str, question_type="tapvid3d_total_distance", img_output_dir=None,
                   npoints_per_group=5, npairs_per_bin=1e8, augment=True, augment_ratio=1.0, max_samples=-1, templates=None,
                   seed=0, device="cuda:0", dot=False):
    """generate_qa_training_data (OM-coord:580-610) for this rank's videos: every video draws from its own
    ``random.Random(seed + index)`` (the reference's Pool workers each inherit the module-level seed), samples are
    concatenated in file order, optionally sub-sampled to ``max_samples``, written as JSONL."""
    out = []
    for i, path in enumerate(npz_files):
        vid_id = os.path.splitext(os.path.basename(path))[0]
        out.extend(generate_video_samples(load_video(path), vid_id, question_type, npoints_per_group, npairs_per_bin,
                                          img_output_dir, None, augment, augment_ratio, templates,
                                          random.Random(seed + i), device, dot))
    if max_samples > 0 and len(out) > max_samples:
        out = random.Random(seed).sample(out, max_samples)
    return streams.write_jsonl(out, output_file)


# ------------------------------------------------------------------------------------------------
# the reference's engine classes (same names, constructor, methods and file layout)
# ------------------------------------------------------------------------------------------------
class TwoFrameVideoQAEngine:
    """single_object_movement_engine_coord.py:281-645.  The reference maps ``generate_qa_training_single_scene`` over a
    ``multiprocessing.Pool`` (OM-coord:584); here the videos of a call run back to back on this process's GPU — give every
    rank its share of ``scene_id_list`` (shard.my_shard) to use several GPUs."""
    dot = False

    def __init__(self, question_type, sub_dataset, templates=None, device="cuda:0"):
        self.question_type, self.sub_dataset, self.templates, self.device = question_type, sub_dataset, templates, device
        loose = self.dot and sub_dataset == "drivetrack"                         # OM-dot:301-303
        self.object_not_moving_threshold = 0.1 if loose else OBJECT_NOT_MOVING_THRESHOLD
        self.camera_not_moving_threshold = 0.1 if loose else CAMERA_NOT_MOVING_THRESHOLD
        self.future_frame_windows = 5 if loose else None                         # OM-dot:303 (1e8 elsewhere: never fires)

    def project_point(self, point_3d, intrinsics, image_height, image_width, id=""):
        """OM-coord:293-315 through msp_track_project (float64 arithmetic on a float32 track position, which is what the
        TAPVid-3D files hold): [u_normalized, v_normalized], or None when the point is outside the image or behind."""
        p = np.asarray(point_3d, dtype=np.float64).reshape(3)
        p32 = p.astype(np.float32)
        if not np.array_equal(p32.astype(np.float64), p):
            raise ValueError("project_point expects a float32 track position (tracks_XYZ); got a float64 value that is not "
                             "float32-representable")
        vid = DeviceVideo(p32.reshape(1, 1, 3), np.eye(4)[None], intrinsics, (image_height, image_width), device=self.device)
        vid.project()
        if not bool(vid.valid.cpu().numpy()[0, 0]):
            u, v = vid.uvn.cpu().numpy()[0, 0]
            print(f"({u}, {v}, {p[2]}) is invaid. Points3D are {point_3d}. intrinsics is {intrinsics}.")
            return None
        return [float(x) for x in vid.uvn.cpu().numpy()[0, 0]]

    def generate_qa_training_single_scene(self, input_file, *args, **kwargs):
        """OM-coord:406 ``(input_file, npoints_per_group=5, npairs_per_bin=1e8, img_output_dir="", augment=True,
        augment_ratio=1.0)``; the dot engine takes ``base_img_dir`` as second positional argument (OM-dot:448)."""
        names = (("base_img_dir",) if self.dot else ()) + ("npoints_per_group", "npairs_per_bin", "img_output_dir", "augment", "augment_ratio")
        a = {"npoints_per_group": 5, "npairs_per_bin": 1e8, "img_output_dir": "", "augment": True, "augment_ratio": 1.0,
             **dict(zip(names, args)), **kwargs}
        video_id = os.path.splitext(os.path.basename(input_file))[0]
        return generate_video_samples(load_video(input_file), video_id, self.question_type, a["npoints_per_group"],
                                      a["npairs_per_bin"], a["img_output_dir"] or None, None, a["augment"], a["augment_ratio"],
                                      self.templates, random, self.device, self.dot, self.object_not_moving_threshold,
                                      self.camera_not_moving_threshold, self.future_frame_windows)

    def _all_videos(self, scene_id_list, source_data_root, npoints_per_group, npairs_per_bin, img_output_dir, augment, augment_ratio):
        data = []
        for scene_id in scene_id_list:
            data.extend(self.generate_qa_training_single_scene(os.path.join(source_data_root, f"{scene_id}.npz"),
                                                               npoints_per_group=npoints_per_group, npairs_per_bin=npairs_per_bin,
                                                               img_output_dir=img_output_dir, augment=augment, augment_ratio=augment_ratio))
        return data

    @staticmethod
    def _report(kind, path, data):
        still = sum(1 for e in data if e["point_moving"] == 0)
        cam_still = sum(1 for e in data if e["cam_moving"] == 0)
        print(f"{kind} data saved to {path}. In total, there are {len(data)} samples.")
        print(f"Object not moving: {still}, Object moving: {len(data) - still}")
        print(f"Camera not moving: {cam_still}, Camera moving: {len(data) - cam_still}")

    def generate_qa_training_data(self, scene_id_list, source_data_root, output_dir, output_file, img_output_dir, npoints_per_group,
                                  npairs_per_bin, augment, augment_ratio=1.0, max_samples=-1, num_workers=20, base_img_dir=None):
        """OM-coord:580-607 / OM-dot:626-652: all videos, optional random.sample, random.shuffle, JSONL."""
        data = self._all_videos(scene_id_list, source_data_root, npoints_per_group, npairs_per_bin, img_output_dir, augment, augment_ratio)
        if max_samples > 0 and len(data) > max_samples:
            data = random.sample(data, max_samples)
        random.shuffle(data)
        streams.write_jsonl(data, output_file)
        self._report("Training", output_file, data)
        return data

    @staticmethod
    def format_eval_sample(training_sample):
        training_sample["text"] = training_sample["conversations"][0]["value"]
        return training_sample

    def generate_qa_eval_data(self, scene_id_list, source_data_root, output_dir, output_file, img_output_dir, npoints_per_group,
                              npairs_per_bin, augment, augment_ratio=0.3, max_samples=300, num_workers=20, base_img_dir=None):
        """OM-coord:614-644 (sub-sample, then eval format) / OM-dot:659-707 (eval format, ``*_orig.jsonl`` with everything, then
        the sub-sample; the reference leaves the sub-sample undefined when nothing has to be dropped — all samples here)."""
        data = self._all_videos(scene_id_list, source_data_root, npoints_per_group, npairs_per_bin, img_output_dir, augment, augment_ratio)
        if self.dot:
            data = [self.format_eval_sample(s) for s in data]
            streams.write_jsonl(data, output_file.replace(".jsonl", "_orig.jsonl"))
            self._report("Original evaluation", output_file.replace(".jsonl", "_orig.jsonl"), data)
        if max_samples > 0 and len(data) > max_samples:
            data = random.sample(data, max_samples)
        if not self.dot:
            data = [self.format_eval_sample(s) for s in data]
        streams.write_jsonl(data, output_file)
        self._report("Subsampled evaluation" if self.dot else "Evaluation", output_file, data)
        return data


class TwoFrameVideoQAEngineDot(TwoFrameVideoQAEngine):
    """single_object_movement_engine_dot.py:294-707 (first image annotated on the device, ids ending in ``_ann``)."""
    dot = True


def main(argv=None):
    """``python -m multi_spatialmllm_b200.tapvid {coord,dot}`` = the ``__main__`` of OM-coord:647-706 / OM-dot:709-779:
    val then train for both question types over adt and pstudio, the reference's directory and file names."""
    import argparse
    import shutil
    ap = argparse.ArgumentParser()
    ap.add_argument("variant", choices=["coord", "dot"])
    ap.add_argument("--base_npz_file", default="data/tapvid3d_dataset")
    ap.add_argument("--meta_data_dir", default="data/tapvid3d_dataset/meta_data")
    ap.add_argument("--train_root", default=None)
    ap.add_argument("--val_root", default=None)
    ap.add_argument("--img_output_dir", default="data/my_tapvid3d_images")
    ap.add_argument("--sub_datasets", default="adt,pstudio")
    ap.add_argument("--version", default="v1_0")
    ap.add_argument("--reference_root", default=None)
    ap.add_argument("--seed", type=int, default=None, help="default: what the reference module sets at import")
    a = ap.parse_args(argv)
    dot = a.variant == "dot"
    train_dir = os.path.join(a.train_root or ("training_data_v2/object_movement_dot" if dot else "training_data/object_movement_coord"), a.version)
    val_dir = os.path.join(a.val_root or ("evaluation_data_v2/object_movement_dot" if dot else "evaluation_data/object_movement_coord"), a.version)
    tp = streams.Templates.from_reference(a.reference_root) if a.reference_root else None
    if a.seed is not None:
        random.seed(a.seed)
    cls = TwoFrameVideoQAEngineDot if dot else TwoFrameVideoQAEngine

    def ids(sub, split):
        with open(os.path.join(a.meta_data_dir, sub, f"{split}.txt")) as f:
            return [ln.rstrip("\n") for ln in f if ln.strip()]

    q_types = ("tapvid3d_total_distance", "tapvid3d_displacement_vector")
    for split, out_root in (("val", val_dir), ("train", train_dir)):
        for q_type in q_types:
            for sub in a.sub_datasets.split(","):
                out = os.path.join(out_root, sub)
                os.makedirs(out, exist_ok=True)
                eng = cls(q_type, sub, templates=tp)
                src = os.path.join(a.base_npz_file, sub)
                if split == "val":
                    img = os.path.join(out, f"{q_type}_images") if dot else a.img_output_dir
                    if dot and os.path.exists(img):
                        shutil.rmtree(img)
                    eng.generate_qa_eval_data(ids(sub, "val"), src, out, os.path.join(out, f"{sub}_{q_type}_val.jsonl"), img, 1, 1,
                                              augment=False, max_samples=300)
                else:
                    img = os.path.join(out, f"{q_type}_images_15points_30pairs") if dot else a.img_output_dir
                    eng.generate_qa_training_data(ids(sub, "train"), src, out, os.path.join(out, f"{sub}_{q_type}_train_15points_30pairs.jsonl"),
                                                  img, 15, 30, augment=True, augment_ratio=0.05)


if __name__ == "__main__":
    main()
