"""Drop-in for spatial_engine/camera_movement/calculate_frames_relations.py — same function names and arguments.

  extract_yaw_pitch(R)                                               CFR:86-100   (msp_pose_prepare)
  calculate_camera_overlap(in_bounds_dict, image_id1, image_id2, use_cuda=False)   CFR:102-137  (msp_pair_overlap)
  process_scene(scene_id, scene_infos, warning_file)                 CFR:139-197
  save_overlap_info / save_overlap_info_nonzero                      CFR:28-84
  run_split(scene_info_path, output_parquet, warning_file, num_workers=15, save_interval=20)   CFR:200-253
  main()                                                             CFR:255-287

``process_scene`` is one K1 launch (every frame's mask) and one ``msp_pair_annotate`` launch (every i<j pair: popcounts of
the two bit rows, pose deltas) instead of two NumPy passes over N bools per pair.  ``calculate_camera_overlap`` keeps the
reference's call shape for callers that hold boolean masks; the masks are packed to bits and counted on the device.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine, tables
from ._lib import check, lib
from .engine import _ptr, _require_cuda, _stream
from .handler.info_handler import SceneInfoHandler


def extract_yaw_pitch(R):
    """(yaw, pitch) in degrees of the camera's z axis; ``R`` 4x4 or 3x3 (CFR:86-100)."""
    _require_cuda()
    R = np.asarray(R, dtype=np.float64)
    M = np.eye(4)
    M[:3, :3] = R[:3, :3]
    dev = torch.device("cuda", torch.cuda.current_device())
    c2w = torch.from_numpy(M.reshape(1, 16)).to(dev)
    E, w2c = torch.empty((1, 16), dtype=torch.float64, device=dev), torch.empty((1, 16), dtype=torch.float64, device=dev)
    yp, valid = torch.empty((1, 2), dtype=torch.float64, device=dev), torch.empty((1,), dtype=torch.int32, device=dev)
    check(lib.msp_pose_prepare(_ptr(c2w), None, 1, _ptr(E), _ptr(w2c), _ptr(yp), _ptr(valid), _stream()), "msp_pose_prepare")
    yaw, pitch = yp.cpu().numpy()[0]
    return np.float64(yaw), np.float64(pitch)


def calculate_camera_overlap(in_bounds_dict, image_id1, image_id2, use_cuda=False):
    """|A & B| / |A | B| * 100 of two boolean visibility masks; NaN (0/0) when neither frame sees a point (CFR:102-137).
    ``use_cuda`` is accepted and ignored: the count always runs on the device."""
    _require_cuda()
    a, b = (np.ascontiguousarray(in_bounds_dict[i], dtype=bool) for i in (image_id1, image_id2))
    if a.shape != b.shape or a.ndim != 1:
        raise ValueError("masks must be 1-D and of equal length")
    n = a.shape[0]
    stride = engine.mask_stride_words(n)
    bits = np.zeros((2, stride * 4), np.uint8)
    for row, m in enumerate((a, b)):
        packed = np.packbits(m, bitorder="little")
        bits[row, :len(packed)] = packed
    dev = torch.device("cuda", torch.cuda.current_device())
    mask = torch.from_numpy(bits.view(np.int32)).to(dev)
    pairs = torch.tensor([[0, 1]], dtype=torch.int32, device=dev)
    inter, uni = (torch.empty((1,), dtype=torch.int32, device=dev) for _ in range(2))
    ov = torch.empty((1,), dtype=torch.float64, device=dev)
    check(lib.msp_pair_overlap(_ptr(mask), n, stride, _ptr(pairs), 1, _ptr(inter), _ptr(uni), _ptr(ov), _stream()),
          "msp_pair_overlap")
    return np.float64(ov.cpu().numpy()[0])


def process_scene(scene_id, scene_infos: SceneInfoHandler, warning_file, device="cuda:0"):
    print(f"Start processing {scene_id}.")
    sc, image_ids = tables.device_scene_from_handler(scene_infos, scene_id, device)
    info = engine.process_scene_relations(sc, image_ids)
    valid = sc.valid.cpu().numpy().astype(bool)
    seen = sc.mask.ne(0).any(dim=1).cpu().numpy()
    with open(warning_file, "a") as f:
        for row, image_id in enumerate(image_ids):                               # CFR:159-161
            if valid[row] and not seen[row]:
                f.write(f"{scene_id}: {image_id} has no in bound points\n")
        for pair, v in info.items():                                             # CFR:191-194
            vals = [np.float64(x) for x in v.values()]                            # printed as the installed NumPy prints them
            if not np.all(np.isfinite(vals)):
                f.write(f"{scene_id}: {pair} has something wrong {vals}. \n")
    print(f"Finished scene {scene_id}.")
    return scene_id, info


def save_overlap_info(overlap_info, parquet_file):
    df = tables.overlap_dataframe(overlap_info)
    if len(df) == 0:
        print(f"[save_overlap_info] Nothing to save to {parquet_file}.")
        return
    df.to_parquet(parquet_file, index=False)
    print(f"[save_overlap_info] Saved {len(df)} records to {parquet_file}.")


def save_overlap_info_nonzero(overlap_info, parquet_file_nonzero):
    df = tables.overlap_dataframe(overlap_info, nonzero_only=True)
    if len(df) == 0:
        print("[save_overlap_info_nonzero] No nonzero-overlap pairs to save.")
        return
    df.to_parquet(parquet_file_nonzero, index=False)
    print(f"[save_overlap_info_nonzero] Saved {len(df)} records to {parquet_file_nonzero}.")


def run_split(scene_info_path, output_parquet, warning_file, num_workers=15, save_interval=20, device="cuda:0"):
    """``num_workers`` / ``save_interval`` are accepted for signature compatibility: one process drives one GPU and the
    file is streamed one row group per scene (``torchrun`` + ``python -m multi_spatialmllm_b200.tables relations`` for more)."""
    scene_infos = scene_info_path if isinstance(scene_info_path, SceneInfoHandler) else SceneInfoHandler(scene_info_path)
    os.makedirs(os.path.dirname(output_parquet) or ".", exist_ok=True)
    res = tables.run_relations_split(scene_infos, output_parquet, device=device, warning_file=warning_file)
    print(f"[run_split] Final save to {output_parquet} complete.")
    print(f"[run_split] Total number of records: {res['rows']}")
    print(f"[run_split] Total number of nonzero records: {res['rows_nonzero']}")
    return res


def main(argv=None):
    return tables.main(["relations"] + list(argv or []))


if __name__ == "__main__":
    import sys
    main(sys.argv[1:])
