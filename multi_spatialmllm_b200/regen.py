"""End-to-end regeneration driver: tables + sample streams for a set of scenes, sharded over GPUs.

    python -m multi_spatialmllm_b200.regen --synthetic 8 --out /tmp/multispa            # seeded synthetic scenes
    torchrun --nproc-per-node 8 -m multi_spatialmllm_b200.regen --info scenes_train_info_i_D5.pkl --out training_data

Per scene (one process per GPU, scenes dealt by ``shard.partition_lpt``; no collective on the data path):
  1. upload the scene, project every frame once (K1 + fused object counts)        make_visibility_info.py / compute_object_visibility.py
  2. all-pairs relations (K3)                                                      calculate_frames_relations.py
  3. sample pairs by overlap bin with the exact replica of the reference's ``sample_dataframe``
     (``pair_engines.sample_rows_by_overlap``: bins [6,35], step 1, quota hand-over from small to large bins;
     camera_movement_engine_train_val.py:29-151 / visual_correspondence_qa_engine_coor_2_coor.py:122-244), frames and
     points for the depth streams (DE-coor:175-199, DC-coor:240-262)
  4. annotate in batch (K4) and write JSONL parts in the reference schemas (streams.py)
Rank 0 gathers the per-rank manifests (shard.gather_manifest) and writes ``manifest.json``.
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np

from . import engine, shard, streams, synth, tables


def regenerate_scene(sc: engine.DeviceScene, scene_id: str, image_ids, n_pair_samples: int, n_depth_frames: int, rng):
    """All streams of one scene; returns {stream: [samples]} and the tables."""
    import torch
    _, counts = sc.project_visibility_objects() if sc.instance_ids is not None and sc.num_objects else (sc.project_visibility(), None)
    valid = np.flatnonzero(sc.valid.cpu().numpy())
    ii, jj = np.triu_indices(len(valid), k=1)
    pairs = np.stack([valid[ii], valid[jj]], 1).astype(np.int32)
    out = {k: [] for k in ("visual_correspondence_coor_2_coor", "camera_movement", "depth_estimation_coor", "depth_comparison_coor")}
    relations = {}
    if len(pairs):
        inter, uni, ov = (t.cpu().numpy() for t in sc.pair_overlap(pairs))
        from .pair_engines import sample_rows_by_overlap
        sel = sample_rows_by_overlap(ov, n_pair_samples, 0, overlap_min=6, overlap_max=35, interval=1, np_random=rng,
                                     log=lambda *a, **k: None)
        chosen = pairs[sel]
        swap = rng.random(len(chosen)) < 0.5                  # random question direction (VC-coor:280-281, CME:163-167)
        chosen[swap] = chosen[swap][:, ::-1]
        words = torch.from_numpy(rng.integers(0, 2**32, len(chosen), dtype=np.uint64).astype(np.uint32).view(np.int32))
        out["visual_correspondence_coor_2_coor"] = streams.visual_correspondence_samples(sc, scene_id, image_ids, chosen, words, 1, rng=rng)
        for qt in streams.CAMERA_QUESTION_TYPES:
            out["camera_movement"] += streams.camera_movement_samples(sc, scene_id, image_ids, chosen, qt, rng=rng)
        relations = {"pairs": pairs, "overlap": ov, "inter": inter, "uni": uni}
    # depth streams: sample frames, then visible points of each frame from the bit masks
    frames = rng.choice(valid, min(n_depth_frames, len(valid)), replace=False) if len(valid) else np.empty(0, np.int64)
    mb = sc.mask_bool()
    de_f, de_p, dc_f, dc_p = [], [], [], []
    for f in frames:
        vis = np.flatnonzero(mb[f])
        if len(vis) >= 1:
            de_f.append(int(f)), de_p.append(int(rng.choice(vis)))
        if len(vis) >= 2:
            dc_f.append(int(f)), dc_p.append(rng.choice(vis, 2, replace=False).tolist())
    if de_f:
        out["depth_estimation_coor"] = streams.depth_estimation_samples(sc, scene_id, image_ids, de_f, de_p, rng=rng)
    if dc_f:
        out["depth_comparison_coor"] = streams.depth_comparison_samples(sc, scene_id, image_ids, dc_f, dc_p,
                                                                       rng.random(len(dc_f)) < 0.5, rng=rng)
    return out, {"relations": relations, "object_counts": None if counts is None else counts.cpu().numpy(), "masks": mb}


def main(argv=None):
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", type=int, default=0, help="number of seeded synthetic scenes (no dataset needed)")
    ap.add_argument("--info", default=None, help="scene-info pkl in the reference layout (real data)")
    ap.add_argument("--posed-images-root", default="data/scannet/posed_images")
    ap.add_argument("--instance-data-root", default="data/scannet/scannet_instance_data")
    ap.add_argument("--out", required=True)
    ap.add_argument("--points", type=int, default=100_000)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--pairs-per-scene", type=int, default=200)
    ap.add_argument("--depth-frames-per-scene", type=int, default=32)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("gloo")                      # host-side gather only; nothing crosses NVLink
    dev = f"cuda:{local}"
    if torch.cuda.is_available():                            # (DeviceScene raises without CUDA: no CPU fallback)
        torch.cuda.set_device(local)                         # libmsp launches on the CURRENT device
    handler = None
    if args.info:
        from .handler.info_handler import SceneInfoHandler
        handler = SceneInfoHandler(args.info, args.posed_images_root, args.instance_data_root)
        costs = {sid: shard.scene_cost(tables._scene_point_count(handler, sid), handler.get_num_posed_images(sid))
                 for sid in handler.get_all_scene_ids()}
    else:
        costs = {f"scene{i:04d}_00": shard.scene_cost(args.points, args.frames) for i in range(args.synthetic)}
    mine = shard.my_shard(costs, rank, world)
    rng = np.random.default_rng(1000 + rank)
    acc, t0 = {}, time.time()
    n_pairs_done = n_frames_done = 0
    for sid in mine:
        if handler is not None:
            sc, image_ids = tables.device_scene_from_handler(handler, sid, dev)
        else:
            s = synth.make_scene(seed=int(sid[5:9]), n_points=args.points, n_frames=args.frames, scene_id=sid)
            sc = engine.DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids,
                                    num_objects=s.num_objects, device=dev)
            sc.render_depth(1)
            image_ids = s.image_ids
        out, tabs = regenerate_scene(sc, sid, image_ids, args.pairs_per_scene, args.depth_frames_per_scene, rng)
        n_frames_done += sc.F
        n_pairs_done += 0 if not tabs["relations"] else len(tabs["relations"]["pairs"])
        for k, v in out.items():
            acc.setdefault(k, []).extend(v)
    torch.cuda.synchronize()
    parts = {}
    for k, v in acc.items():
        path = shard.part_path(os.path.join(args.out, f"{k}.jsonl"), rank)
        parts[k] = {"path": path, "samples": streams.write_jsonl(v, path)}
    secs = max(time.time() - t0, 1e-9)
    n_samples = sum(p["samples"] for p in parts.values())
    man = shard.gather_manifest({"rank": rank, "scenes": mine, "seconds": round(secs, 2), "parts": parts, "samples": n_samples,
                                 "frames": n_frames_done, "pairs_related": n_pairs_done,
                                 "pairs_per_s": round(n_pairs_done / secs, 1), "frames_per_s": round(n_frames_done / secs, 1),
                                 "samples_per_s": round(n_samples / secs, 1)},
                                os.path.join(args.out, "manifest.json"))
    print(f"[regen rank {rank}] {len(mine)} scenes, {n_frames_done} frames, {n_pairs_done} pairs related, {n_samples} samples in "
          f"{secs:.1f} s: {n_pairs_done / secs:.0f} pairs/s, {n_samples / secs:.0f} samples/s (JSONL written)", flush=True)
    if rank == 0:
        for k in acc:
            shard.merge_jsonl_parts([m["parts"][k]["path"] for m in man if k in m["parts"]], os.path.join(args.out, f"{k}.jsonl"))
        print(f"[regen] {sum(m['samples'] for m in man)} samples from {sum(len(m['scenes']) for m in man)} scenes -> {args.out}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
