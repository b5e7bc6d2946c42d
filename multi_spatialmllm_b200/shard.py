"""Scene / video sharding across GPUs: one process per GPU, no data-path collective.

The reference parallelises by scene with multiprocessing.Pool (make_visibility_info.py:151,
calculate_frames_relations.py:222, single_object_movement_engine_coord.py:584); scenes share no state.  Here each
rank owns a cost-balanced subset end to end and writes its own part file; the only cross-rank step is one
host-side gather of the per-rank manifests (torch.distributed gather_object — gloo on CPU, nccl-backed groups
fall back to their object collectives), after which rank 0 writes the merged manifest.
"""
from __future__ import annotations

import json
import os


def scene_cost(n_points: int, n_frames: int, n_pairs: int | None = None) -> float:
    """Work estimate of SURVEY.md §8(e): projection N*F plus the pair stage P*N/8 (all pairs when P is None)."""
    p = n_frames * (n_frames - 1) // 2 if n_pairs is None else n_pairs
    return float(n_points) * n_frames + p * n_points / 8.0


def partition_lpt(costs: dict, world_size: int) -> list:
    """Greedy longest-processing-time assignment: ids sorted by cost, each to the least-loaded rank.
    Deterministic (ties broken by id) so every rank computes the same partition without communicating."""
    bins = [[] for _ in range(world_size)]
    load = [0.0] * world_size
    for sid in sorted(costs, key=lambda s: (-costs[s], s)):
        r = min(range(world_size), key=lambda k: (load[k], k))
        bins[r].append(sid)
        load[r] += costs[sid]
    return bins


def my_shard(costs: dict, rank: int, world_size: int) -> list:
    return partition_lpt(costs, world_size)[rank]


def part_path(output_file: str, rank: int) -> str:
    root, ext = os.path.splitext(output_file)
    return f"{root}.part-{rank:05d}{ext}"


def gather_manifest(local_manifest: dict, manifest_file: str | None = None):
    """One host-side gather of per-rank manifests; rank 0 returns (and optionally writes) the merged list."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        merged = [local_manifest]
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = [None] * world if rank == 0 else None
        dist.gather_object(local_manifest, buf, dst=0)
        if rank != 0:
            return None
        merged = buf
    if manifest_file:
        os.makedirs(os.path.dirname(manifest_file) or ".", exist_ok=True)
        with open(manifest_file, "w") as f:
            json.dump({"parts": merged, "total_samples": sum(m.get("samples", 0) for m in merged)}, f, indent=1)
    return merged


def merge_jsonl_parts(part_files, output_file: str) -> int:
    n = 0
    with open(output_file, "w") as out:
        for p in part_files:
            with open(p) as f:
                for ln in f:
                    out.write(ln)
                    n += 1
    return n


def merge_text_parts(part_files, output_file: str) -> int:
    """Append the per-rank text parts (e.g. warning files of ranks > 0) to rank 0's file, in rank order; missing parts
    are skipped and merged parts removed.  Returns the number of lines appended."""
    n = 0
    with open(output_file, "a") as out:
        for p in part_files:
            if not os.path.exists(p):
                continue
            with open(p) as f:
                for ln in f:
                    out.write(ln)
                    n += 1
            os.remove(p)
    return n
