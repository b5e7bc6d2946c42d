"""Videos per second of the TAPVid-3D object_movement worker (BASELINE config 3 shape: T = 300 frames, N = 1024 tracks,
360x640 frames) with the reference's training settings (15 points per group, 30 pairs per bin, augment 0.05).

    python profiles/tapvid_bench.py ours        (GPU box)  tapvid.generate_video_samples per clip, frames exported
    python profiles/tapvid_bench.py reference   (build container, needs /root/reference) the unmodified
                                                generate_qa_training_single_scene on the same clips, one process
"""
import contextlib
import json
import os
import random
import sys
import tempfile
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_spatialmllm_b200 import synth  # noqa: E402

T, N, CLIPS = int(os.environ.get("TB_T", 300)), int(os.environ.get("TB_N", 1024)), int(os.environ.get("TB_CLIPS", 4))
HW = (360, 640)


def write_clips(tmp):
    ok, jpg = cv2.imencode(".jpg", np.full((*HW, 3), 90, np.uint8))
    files = []
    for k in range(CLIPS):
        v = synth.make_video(200 + k, n_frames=T, n_tracks=N, n_groups=6, image_hw=HW)
        p = os.path.join(tmp, f"{v.video_id}.npz")
        np.savez(p, images_jpeg_bytes=np.array([jpg.tobytes()] * T, dtype=object), fx_fy_cx_cy=v.fxfycxcy,
                 tracks_XYZ=v.tracks_xyz, visibility=v.visibility, queries_xyt=np.zeros((N, 3)), extrinsics_w2c=v.w2c)
        files.append(p)
    return files


def main(side):
    quiet = contextlib.redirect_stdout(open(os.devnull, "w"))
    with tempfile.TemporaryDirectory() as tmp:
        files = write_clips(tmp)
        out = {"side": side, "frames": T, "tracks": N, "clips": CLIPS, "per_clip_s": [], "samples": 0}
        if side == "ours":
            import torch
            from multi_spatialmllm_b200 import tapvid
            eng = tapvid.TwoFrameVideoQAEngine("tapvid3d_total_distance", "adt")
            out["gpu"] = torch.cuda.get_device_name(0)
            random.seed(0)
            eng.generate_qa_training_single_scene(files[0], 15, 30, os.path.join(tmp, "warm"), True, 0.05)   # warm-up
            fn = lambda f: eng.generate_qa_training_single_scene(f, 15, 30, os.path.join(tmp, "img"), True, 0.05)
        else:
            from oracle import ref_shim
            ref_shim.import_reference()
            with quiet:
                from spatial_engine.object_movement import single_object_movement_engine_coord as om
            eng = om.TwoFrameVideoQAEngine(question_type="tapvid3d_total_distance", sub_dataset="adt")
            out["cores_used"], out["host_cores"] = 1, os.cpu_count()
            fn = lambda f: eng.generate_qa_training_single_scene(f, 15, 30, os.path.join(tmp, "img"), True, 0.05)
            files = files[:int(os.environ.get("TB_REF_CLIPS", 2))]
        random.seed(0)
        for f in files:
            t = time.time()
            with quiet, np.errstate(all="ignore"):
                s = fn(f)
            out["per_clip_s"].append(round(time.time() - t, 3))
            out["samples"] += len(s)
        out["videos_per_s"] = round(len(files) / sum(out["per_clip_s"]), 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "ours")
