"""L2 tables in the reference's on-disk schemas + the per-split drivers that fill them.

  visibility table  parquet, columns key / values; key "{scene}:image_to_points:{image_id}" or
                    "{scene}:point_to_images:{point}", values = JSON string (make_visibility_info.py:59-65 —
                    the canonical form every reader expects, info_handler.py:510-513; the keys written by
                    run_split at make_visibility_info.py:163-172 are not readable by the reference itself)
  overlap table     parquet, columns scene_id,image_id1,image_id2,overlap,distance,yaw,pitch
                    (calculate_frames_relations.py:43-51) + the "_nonzero" companion (:59-84)
  object visibility pkl {scene: {"object_to_images": {...}, "image_to_objects": {...}}}
                    (compute_object_visibility.py:14-38,139-150)

Geometry comes from engine.DeviceScene (CUDA); this module is host-side formatting and file I/O only.
"""
from __future__ import annotations

import json
import os
import pickle

import numpy as np
import pandas as pd

from . import engine

NONINFORMATIVE_DESC = {"wall", "object", "floor", "ceiling", "window"}     # compute_object_visibility.py:58


def visibility_rows(scene_id: str, table: dict):
    for image_id, pts in table["image_to_points"].items():
        yield f"{scene_id}:image_to_points:{image_id}", json.dumps(pts)
    for point, images in table["point_to_images"].items():
        yield f"{scene_id}:point_to_images:{point}", json.dumps(images)


def save_visibility_parquet(scene_tables: dict, parquet_file: str) -> int:
    rows = [r for sid, t in scene_tables.items() for r in visibility_rows(sid, t)]
    os.makedirs(os.path.dirname(parquet_file) or ".", exist_ok=True)
    pd.DataFrame(rows, columns=["key", "values"]).to_parquet(parquet_file, index=False)
    return len(rows)


def overlap_dataframe(overlap_info: dict, nonzero_only: bool = False) -> pd.DataFrame:
    cols = {k: [] for k in ("scene_id", "image_id1", "image_id2", "overlap", "distance", "yaw", "pitch")}
    for sid, pairs in overlap_info.items():
        for (a, b), v in pairs.items():
            if nonzero_only and not (v["overlap"] != 0.0):      # NaN != 0.0 is True: kept, like CFR:67
                continue
            cols["scene_id"].append(sid), cols["image_id1"].append(a), cols["image_id2"].append(b)
            for k in ("overlap", "distance", "yaw", "pitch"):
                cols[k].append(v[k])
    return pd.DataFrame(cols)


def save_overlap_parquet(overlap_info: dict, parquet_file: str):
    """Writes <file> and <file>_nonzero.parquet like run_split (calculate_frames_relations.py:241-245)."""
    os.makedirs(os.path.dirname(parquet_file) or ".", exist_ok=True)
    df = overlap_dataframe(overlap_info)
    if len(df):
        df.to_parquet(parquet_file, index=False)
    nz = overlap_dataframe(overlap_info, nonzero_only=True)
    if len(nz):
        nz.to_parquet(parquet_file.replace(".parquet", "_nonzero.parquet"), index=False)
    return len(df), len(nz)


def device_scene_from_handler(scene_infos, scene_id: str, device="cuda:0"):
    """Upload one scene of a SceneInfoHandler: aligned points, every frame's pose (invalid ones included, they
    are masked on the device) and its depth PNG.  Returns (DeviceScene, image_ids in row order)."""
    image_ids = scene_infos.get_all_image_ids(scene_id)
    pts = np.ascontiguousarray(scene_infos.get_scene_points_align(scene_id)[:, :3], dtype=np.float64)
    c2w = np.stack([scene_infos.get_extrinsic_matrix(scene_id, i, warning=False) for i in image_ids])
    H, W = scene_infos.get_image_shape(scene_id)
    depth = None
    if len(image_ids) > 8:                                   # cv2.imread releases the GIL: decode the PNGs on a few threads
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(min(8, os.cpu_count() or 1)) as pool:
            maps = list(pool.map(lambda i: scene_infos.get_depth_image(scene_id, i), image_ids))
    else:
        maps = [scene_infos.get_depth_image(scene_id, i) for i in image_ids]
    for f, d in enumerate(maps):
        if d is None:
            continue
        if depth is None:
            depth = np.zeros((len(image_ids), *d.shape[:2]), np.uint16)
        depth[f] = d
    inst = n_obj = None
    try:
        inst = scene_infos.get_scene_instance_mask(scene_id).astype(np.uint16)
        n_obj = int(scene_infos.get_num_objects(scene_id))
    except Exception:
        pass
    sc = engine.DeviceScene(pts, scene_infos.get_intrinsic_matrix(scene_id), c2w,
                            axis_align=scene_infos.get_world_to_axis_align_matrix(scene_id), depth=depth,
                            instance_ids=inst, num_objects=n_obj or 0, image_hw=(H, W), depth_hw=depth.shape[1:],
                            depth_scale=scene_infos.depth_value_scale, device=device)
    return sc, image_ids


# ------------------------------------------------------------------------------------------------
# device-formatted tables: the JSON text / columns leave the GPU ready for Arrow, no Python lists on the way
# ------------------------------------------------------------------------------------------------
_JSON_SAFE = frozenset(chr(c) for c in range(0x20, 0x7f)) - {'"', "\\"}


class _PinnedPool:
    """Grow-only pool of pinned host buffers for the table text (200 MB per 200 k-point scene): the device->host copy of a
    table runs at PCIe speed instead of the ~2 GB/s a fresh pageable allocation gets (page faults), and the buffer goes back
    to the pool when the Arrow table that wraps it (zero-copy) is released — possibly on a writer thread, hence the lock."""

    def __init__(self):
        import threading
        self._free, self._lock = [], threading.Lock()

    def take(self, nbytes: int):
        import torch
        with self._lock:
            fit = [k for k, t in enumerate(self._free) if t.numel() >= nbytes]
            if fit:
                k = min(fit, key=lambda i: self._free[i].numel())
                return self._free.pop(k)            # (list.remove would compare tensors element-wise)
        return torch.empty((max(int(nbytes * 1.25), 1 << 20),), dtype=torch.uint8, pin_memory=True)

    def give(self, t):
        with self._lock:
            if len(self._free) < 24:
                self._free.append(t)


_TEXT_POOL = _PinnedPool()


def _to_host_pooled(dev_bytes):
    """uint8 device tensor -> NumPy view of a pooled pinned buffer; the buffer returns to the pool when the view dies."""
    import weakref
    n = int(dev_bytes.numel())
    pin = _TEXT_POOL.take(n)
    pin[:n].copy_(dev_bytes)
    arr = pin[:n].numpy()
    weakref.finalize(arr, _TEXT_POOL.give, pin)
    return arr


def _check_json_safe(ids):
    for s in ids:
        if not isinstance(s, str) or len(s) > 255 or not set(s) <= _JSON_SAFE:
            raise ValueError(f"frame id {s!r} would be escaped by json.dumps (or is longer than 255 bytes): not supported "
                             "by the device formatter")


def visibility_arrow_table(sc: "engine.DeviceScene", scene_id: str, image_ids):
    """The rows convert_pkl_to_parquet writes for one scene (MVI:59-65) as a pyarrow Table (key, values).

    Row order and bytes are the reference's: one ``{scene}:image_to_points:{image_id}`` row per valid frame in
    ``image_ids`` order (MVI:93-104), then one ``{scene}:point_to_images:{p}`` row per point (MVI:116-118, ids sorted as
    strings).  The ``values`` text is written on the device from the bit masks (msp_visibility_json_*); the host only
    wraps the returned buffers.
    """
    import ctypes as C  # noqa: F401
    import pyarrow as pa
    import pyarrow.compute as pc
    import torch
    from ._lib import check, lib
    from .engine import _ptr, _stream
    if sc.perm is not None:
        raise ValueError("the table formatter needs the caller's point order (reorder=None)")
    if sc.mask is None:
        sc.project_visibility()
    dev = sc.device
    valid = sc.valid.cpu().numpy().astype(bool)
    assert len(image_ids) == sc.F, "one id per frame row"
    rows_a = np.flatnonzero(valid).astype(np.int32)
    ids_a = [image_ids[f] for f in rows_a]
    _check_json_safe(ids_a)
    order = np.argsort(np.asarray(ids_a, dtype=object).astype(str), kind="stable") if len(ids_a) else np.empty(0, np.int64)
    rows_b = rows_a[order]
    ids_b = [ids_a[k] for k in order]
    if any(ids_b[k] == ids_b[k + 1] for k in range(len(ids_b) - 1)):
        raise ValueError("duplicate frame ids")                      # the reference's set() would merge them
    R, n = int(len(rows_a)), sc.n
    pitch = max([len(s) for s in ids_b], default=1) or 1
    id_tab = np.zeros((max(R, 1), pitch), np.uint8)
    id_len = np.zeros(max(R, 1), np.uint8)
    for k, s_ in enumerate(ids_b):
        b = s_.encode("ascii")
        id_tab[k, :len(b)] = np.frombuffer(b, np.uint8)
        id_len[k] = len(b)
    rows_a_d = torch.from_numpy(rows_a).to(dev)
    rows_b_d = torch.from_numpy(np.ascontiguousarray(rows_b)).to(dev)
    id_tab_d, id_len_d = torch.from_numpy(id_tab).to(dev), torch.from_numpy(id_len).to(dev)
    sizes = torch.empty(R + n, dtype=torch.int64, device=dev)
    st = _stream()
    check(lib.msp_visibility_json_points_size(_ptr(sc.mask), n, sc.stride, _ptr(rows_a_d), R, sizes.data_ptr(), st),
          "msp_visibility_json_points_size")
    check(lib.msp_visibility_json_images_size(_ptr(sc.mask), n, sc.stride, _ptr(rows_b_d), R, _ptr(id_len_d),
                                              sizes[R:].data_ptr() if n else None, st), "msp_visibility_json_images_size")
    offs = torch.zeros(R + n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(sizes, 0, out=offs[1:])
    total = int(offs[-1].item())                                     # the one sync: the text buffer has to be sized
    text = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    check(lib.msp_visibility_json_points_write(_ptr(sc.mask), n, sc.stride, _ptr(rows_a_d), R, offs.data_ptr(),
                                               text.data_ptr(), st), "msp_visibility_json_points_write")
    check(lib.msp_visibility_json_images_write(_ptr(sc.mask), n, sc.stride, _ptr(rows_b_d), R, _ptr(id_tab_d), pitch,
                                               _ptr(id_len_d), offs[R:].data_ptr(), text.data_ptr(), st),
          "msp_visibility_json_images_write")
    small = total < 2**31
    offs_h = (offs.to(torch.int32) if small else offs).cpu().numpy()
    text_h = _to_host_pooled(text[:total])
    cls = pa.StringArray if small else pa.LargeStringArray
    values = cls.from_buffers(R + n, pa.py_buffer(offs_h), pa.py_buffer(text_h))
    str_t = pa.string() if small else pa.large_string()
    key_a = pa.array([f"{scene_id}:image_to_points:{i}" for i in ids_a], type=str_t)
    key_b = pc.binary_join_element_wise(pa.scalar(f"{scene_id}:point_to_images:", type=str_t),
                                        pc.cast(pa.array(np.arange(n, dtype=np.int64)), str_t), pa.scalar("", type=str_t))
    return pa.table({"key": pa.concat_arrays([key_a, key_b]), "values": values})


def relations_arrow_table(sc: "engine.DeviceScene", scene_id: str, image_ids):
    """calculate_frames_relations.process_scene + save_overlap_info (CFR:139-197, 28-57) for one scene as a pyarrow
    Table with the reference's columns; every i<j pair of valid frames in itertools.combinations order, no Python loop."""
    import pyarrow as pa
    if sc.mask is None:
        sc.project_visibility()
    valid = np.flatnonzero(sc.valid.cpu().numpy())
    ii, jj = np.triu_indices(len(valid), k=1)                        # == combinations(valid, 2) order
    pairs = np.stack([valid[ii], valid[jj]], 1).astype(np.int32)
    P = len(pairs)
    if P:
        rec = sc.records_to_numpy(sc.pair_annotate(pairs, np.zeros(P, np.uint32), sel_mode=0))
    else:
        rec = np.zeros(0, engine.PAIR_RECORD_DTYPE)
    ids = pa.array([str(i) for i in image_ids], type=pa.string())
    return pa.table({
        "scene_id": pa.DictionaryArray.from_arrays(pa.array(np.zeros(P, np.int32)), pa.array([scene_id])).cast(pa.string()),
        "image_id1": ids.take(pa.array(pairs[:, 0])) if P else pa.array([], pa.string()),
        "image_id2": ids.take(pa.array(pairs[:, 1])) if P else pa.array([], pa.string()),
        "overlap": pa.array(np.ascontiguousarray(rec["overlap"])), "distance": pa.array(np.ascontiguousarray(rec["distance"])),
        "yaw": pa.array(np.ascontiguousarray(rec["dyaw"])), "pitch": pa.array(np.ascontiguousarray(rec["dpitch"])),
    })


class _ParquetStream:
    """Row-group-per-scene parquet writer; the file is created by the first non-empty table (the reference writes no file
    when there is nothing to save, CFR:52-54).  Writes run on a background thread so that Arrow's encoding/compression of
    scene k overlaps the GPU work of scene k+1.

    ``writers > 1`` selects the DATASET layout: ``path`` becomes a directory (still named ``*.parquet``) with one complete
    parquet file per scene, ``part-000000.parquet`` ..., encoded by ``writers`` threads in parallel (pyarrow releases the
    GIL).  ``pandas.read_parquet(path)`` / ``pyarrow.parquet.read_table(path)`` — what ``VisibilityInfoHandler`` does,
    IH:451-460 — read such a directory in file-name order, i.e. the same rows in the same order as the single file; the
    snappy encode of a 200 k-point scene's list text (0.3-0.5 s on one core) is what bounds the single-file writer."""

    def __init__(self, path: str, writers: int = 1):
        from concurrent.futures import ThreadPoolExecutor
        self.path, self.writer, self.rows = path, None, 0
        self.writers = max(1, int(writers))
        self._pool, self._pending, self._parts = ThreadPoolExecutor(self.writers), [], 0
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)

    def _write(self, table):
        import pyarrow.parquet as pq
        if self.writer is None:
            self.writer = pq.ParquetWriter(self.path, table.schema)
        self.writer.write_table(table)

    def _write_part(self, table, k):
        import pyarrow.parquet as pq
        pq.write_table(table, os.path.join(self.path, f"part-{k:06d}.parquet"))

    def write(self, table):
        if table.num_rows == 0:
            return
        self.rows += table.num_rows
        if self.writers == 1:
            if self._pending:
                self._pending.pop().result()
            self._pending.append(self._pool.submit(self._write, table))
            return
        if self._parts == 0:
            if os.path.isfile(self.path):
                os.remove(self.path)
            os.makedirs(self.path, exist_ok=True)
            for old in os.listdir(self.path):
                if old.startswith("part-") and old.endswith(".parquet"):
                    os.remove(os.path.join(self.path, old))
        while len(self._pending) >= 2 * self.writers:          # bound the tables held in host memory
            self._pending.pop(0).result()
        self._pending.append(self._pool.submit(self._write_part, table, self._parts))
        self._parts += 1

    def close(self) -> int:
        for f in self._pending:
            f.result()
        self._pending = []
        self._pool.shutdown()
        if self.writer is not None:
            self.writer.close()
        return self.rows


def _warn_empty_frames(sc, scene_id, image_ids, warning_file, text):
    """MVI:110-113 / CFR:159-161: one line per valid frame that sees no point."""
    if not warning_file:
        return
    valid = sc.valid.cpu().numpy().astype(bool)
    seen = sc.mask.ne(0).any(dim=1).cpu().numpy()
    lines = [text.format(scene_id=scene_id, image_id=i) for f, i in enumerate(image_ids) if valid[f] and not seen[f]]
    if lines:
        with open(warning_file, "a") as f:
            f.writelines(lines)


def run_visibility_split(scene_infos, output_file: str, scene_ids=None, device="cuda:0", warning_file=None, writers: int = 1) -> dict:
    """make_visibility_info.run_split + convert_pkl_to_parquet (MVI:127-177, 38-73) on the GPU: streams one row group
    per scene into the canonical parquet (same keys, same JSON text, same row order as the reference's file).
    ``writers > 1``: a ``*.parquet`` dataset directory with one file per scene, encoded on that many host threads (same
    rows for every reader of the path; see _ParquetStream)."""
    out = _ParquetStream(output_file, writers=writers)
    n_scenes = 0
    for sid in (scene_ids if scene_ids is not None else scene_infos.get_all_scene_ids()):
        sc, image_ids = device_scene_from_handler(scene_infos, sid, device)
        out.write(visibility_arrow_table(sc, sid, image_ids))
        _warn_empty_frames(sc, sid, image_ids, warning_file, "[Warning] {scene_id}: {image_id} has no in-bound points.\n")
        n_scenes += 1
    return {"scenes": n_scenes, "rows": out.close(), "file": output_file}


def run_relations_split(scene_infos, output_parquet: str, scene_ids=None, device="cuda:0", warning_file=None) -> dict:
    """calculate_frames_relations.run_split (CFR:200-253) on the GPU: <file> and <file>_nonzero.parquet (CFR:59-84;
    NaN != 0.0, so pairs with an empty union stay in the nonzero table like in the reference)."""
    import pyarrow.compute as pc
    full = _ParquetStream(output_parquet)
    nonzero = _ParquetStream(output_parquet.replace(".parquet", "_nonzero.parquet"))
    n_scenes = 0
    for sid in (scene_ids if scene_ids is not None else scene_infos.get_all_scene_ids()):
        sc, image_ids = device_scene_from_handler(scene_infos, sid, device)
        t = relations_arrow_table(sc, sid, image_ids)
        full.write(t)
        _warn_empty_frames(sc, sid, image_ids, warning_file, "{scene_id}: {image_id} has no in bound points\n")
        if t.num_rows:
            ov = t["overlap"]
            nonzero.write(t.filter(pc.or_(pc.is_nan(ov), pc.not_equal(ov, 0.0))))
            if warning_file:                                                      # CFR:191-194: rows holding NaN / inf
                bad = t.filter(pc.invert(pc.and_(pc.and_(pc.is_finite(t["overlap"]), pc.is_finite(t["distance"])),
                                                 pc.and_(pc.is_finite(t["yaw"]), pc.is_finite(t["pitch"]))))).to_pylist()
                if bad:
                    with open(warning_file, "a") as f:
                        for r in bad:
                            vals = [np.float64(r[k]) for k in ("overlap", "distance", "yaw", "pitch")]
                            f.write(f"{sid}: {(r['image_id1'], r['image_id2'])} has something wrong {vals}. \n")
        n_scenes += 1
    return {"scenes": n_scenes, "rows": full.close(), "rows_nonzero": nonzero.close(), "file": output_parquet}


def object_visibility_scene(scene_infos, scene_id: str, sc: "engine.DeviceScene", image_ids, warnings_list=None) -> dict:
    """compute_object_visibility.process_scene (COV:72-152) from the device's per-(frame, object) counts; the reference's
    warning lines (objects without points, COV:112-116) are appended to ``warnings_list`` when given."""
    result = {"object_to_images": {}, "image_to_objects": {}}
    sc.project_visibility()
    counts = sc.object_counts().cpu().numpy()
    valid = sc.valid.cpu().numpy().astype(bool)
    inst = scene_infos.get_scene_instance_mask(scene_id)
    for o in range(scene_infos.get_num_objects(scene_id)):
        if scene_infos.get_object_raw_category(scene_id, o) in NONINFORMATIVE_DESC:
            continue
        total = int(np.count_nonzero(inst == o + 1))
        if total == 0:
            if warnings_list is not None:
                warnings_list.append(f"[Warning] Scene {scene_id}, object {o} has no point indices, skipping.")
                print(warnings_list[-1])
            continue
        thr = max(1, int(0.05 * total))
        for f, image_id in enumerate(image_ids):
            c = int(counts[f, o])
            if valid[f] and c >= thr:
                row = {"intersection_count": c, "visibility": (c / total) * 100.0}
                result["object_to_images"].setdefault(o, []).append({"image_id": image_id, **row})
                result["image_to_objects"].setdefault(image_id, []).append({"object_id": o, **row})
    return result


def run_object_visibility_split(scene_infos, output_pkl: str, scene_ids=None, device="cuda:0", warning_file=None) -> dict:
    """compute_object_visibility.process_split (COV:154-194): {scene: result} pickled to ``output_pkl``."""
    out, warnings_list = {}, []
    for sid in (scene_ids if scene_ids is not None else scene_infos.get_all_scene_ids()):
        sc, image_ids = device_scene_from_handler(scene_infos, sid, device)
        out[sid] = object_visibility_scene(scene_infos, sid, sc, image_ids, warnings_list)
    if warning_file:
        with open(warning_file, "a") as wf:
            wf.writelines(w + "\n" for w in warnings_list)
    os.makedirs(os.path.dirname(output_pkl) or ".", exist_ok=True)
    with open(output_pkl, "wb") as f:
        pickle.dump(out, f)
    return out


COVERAGE_TOLERANCE = 0.1        # single_object_coverage_finder.py TOLERANCE


def object_coverage_scene(scene_infos, scene_id: str, sc: "engine.DeviceScene", image_ids, object_visibility: dict,
                          max_images: int = 5) -> dict:
    """single_object_coverage_finder.process_scene_for_coverage (SOCF:267-284) with process_object (SOCF:222-265):
    per object {"height" | "length" | "width": {k: [image-id tuples]}}, using one msp_object_extents pass for the scene
    instead of union masks.  Uses the global ``random`` state like the reference (seed it the same way)."""
    from . import coverage
    sc.project_visibility()
    ext = sc.object_extents().cpu().numpy()                      # (F, n_obj, 6)
    row = {img: f for f, img in enumerate(image_ids)}
    out = {}
    for object_id, vis_list in object_visibility["object_to_images"].items():
        imgs = [v["image_id"] for v in vis_list]
        rows = [row[i] for i in imgs]
        width_axis = scene_infos.get_object_width_axis_aligned(scene_id, object_id)
        plan = (("height", 2, scene_infos.get_object_height(scene_id, object_id)),
                ("length", 1 if width_axis == 0 else 0, scene_infos.get_object_length(scene_id, object_id)),
                ("width", width_axis, scene_infos.get_object_width(scene_id, object_id)))
        res = {}
        for name, axis, target in plan:
            res[name] = coverage.find_minimal_combinations(imgs, ext[rows, object_id, axis], ext[rows, object_id, 3 + axis],
                                                           float(target), COVERAGE_TOLERANCE, max_images=max_images)
        out[object_id] = res
    return out


# ------------------------------------------------------------------------------------------------
# the three L2 scripts' main(): reference default paths, one process per GPU under torchrun
# ------------------------------------------------------------------------------------------------
L2_DEFAULTS = {                                        # MVI:190-206, CFR:257-268, COV:199-210
    "visibility": {"train": "data/scannet/scannet_instance_data/train_visibility_info_D5.parquet",
                   "val": "data/scannet/scannet_instance_data/val_visibility_info_D5.parquet"},
    "relations": {"train": "training_data/camera_movement/train_camera_info_D5.parquet",
                  "val": "evaluation_data/camera_movement/val_camera_info_D5.parquet"},
    "object_visibility": {"train": "training_data/object_perception/object_visibility.pkl",
                          "val": "evaluation_data/object_perception/object_visibility.pkl"},
}
L2_INFO = {"train": "data/scannet/scannet_instance_data/scenes_train_info_i_D5.pkl",
           "val": "data/scannet/scannet_instance_data/scenes_val_info_i_D5.pkl"}


def merge_parquet_parts(part_files, output_file: str, scene_order, scene_of=None) -> int:
    """Concatenate the per-rank part files (one row group per scene) into ``output_file`` with the scenes in
    ``scene_order`` — the row order the single-process run (and the reference) produces.  ``scene_of`` maps the first row
    of a row group to its scene id (default: the ``scene_id`` column, else the ``key`` prefix)."""
    import pyarrow.parquet as pq
    where = {}
    files = [pq.ParquetFile(p) for p in part_files if os.path.exists(p)]
    for pf in files:
        col = "scene_id" if "scene_id" in pf.schema_arrow.names else "key"
        for g in range(pf.num_row_groups):
            first = pf.read_row_group(g, columns=[col]).column(0)[0].as_py()
            sid = scene_of(first) if scene_of else (first if col == "scene_id" else first.split(":", 1)[0])
            where.setdefault(sid, []).append((pf, g))
    rows, writer = 0, None
    for sid in scene_order:
        for pf, g in where.get(sid, []):
            t = pf.read_row_group(g)
            if writer is None:
                writer = pq.ParquetWriter(output_file, t.schema)
            writer.write_table(t)
            rows += t.num_rows
    if writer is not None:
        writer.close()
    return rows


def _scene_point_count(handler, scene_id) -> int:
    """Number of points of a scene for the cost model WITHOUT loading the array: the .npy header (memory-mapped) when
    the handler exposes the file, otherwise the array itself."""
    try:
        path = os.path.join(handler.instance_data_root, scene_id, "aligned_points.npy")
        if not os.path.exists(path):
            path = os.path.join(handler.instance_data_root, f"{scene_id}_aligned_points.npy")
        return int(np.load(path, mmap_mode="r").shape[0])
    except Exception:
        return len(handler.get_scene_points_align(scene_id))


def main(argv=None):
    """``python -m multi_spatialmllm_b200.tables {visibility,relations,object_visibility}`` = the ``main()`` of
    make_visibility_info.py, calculate_frames_relations.py and compute_object_visibility.py (val split, then train, the
    reference's default paths).  Under ``torchrun`` every rank takes a cost-balanced share of the scenes (shard.py), writes
    its own part file, and rank 0 merges the parts in scene order after one host-side gather — no device collective."""
    import argparse
    from . import shard
    from .handler.info_handler import SceneInfoHandler
    ap = argparse.ArgumentParser()
    ap.add_argument("table", choices=sorted(L2_DEFAULTS))
    ap.add_argument("--splits", default="val,train")
    ap.add_argument("--train_info_path", default=L2_INFO["train"])
    ap.add_argument("--val_info_path", default=L2_INFO["val"])
    ap.add_argument("--train_output", default=None)
    ap.add_argument("--val_output", default=None)
    ap.add_argument("--posed_images_root", default="data/scannet/posed_images")
    ap.add_argument("--instance_data_root", default="data/scannet/scannet_instance_data")
    ap.add_argument("--writers", type=int, default=1, help="visibility table, single process: > 1 writes a *.parquet dataset "
                    "directory (one file per scene) with that many encode threads instead of the reference's single file")
    a = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    import torch
    if torch.cuda.is_available():                            # (the scene workers raise without CUDA: no CPU fallback)
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))   # libmsp launches on the CURRENT device
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo")                  # host-side gather only
    run = {"visibility": run_visibility_split, "relations": run_relations_split,
           "object_visibility": run_object_visibility_split}[a.table]
    results = {}
    for split in a.splits.split(","):
        h = SceneInfoHandler(getattr(a, f"{split}_info_path"), posed_images_root=a.posed_images_root,
                             instance_data_root=a.instance_data_root)
        out = getattr(a, f"{split}_output") or L2_DEFAULTS[a.table][split]
        os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
        scene_ids = h.get_all_scene_ids()
        warn = os.path.join(os.path.dirname(out) or ".", {"visibility": f"make_visibility_{split}_warning.txt",      # MVI:200-201
                                                          "relations": f"{split}_warning_D5.txt",                    # CFR:270-271
                                                          "object_visibility": "warning.txt"}[a.table])              # COV:166
        if rank > 0:
            warn = shard.part_path(warn, rank)
        if world == 1:
            kw = {"writers": a.writers} if a.table == "visibility" and a.writers > 1 else {}
            results[split] = run(h, out, device=device, warning_file=warn, **kw)
            continue
        costs = {s: shard.scene_cost(_scene_point_count(h, s), h.get_num_posed_images(s)) for s in scene_ids}
        mine = set(shard.my_shard(costs, rank, world))
        part = shard.part_path(out, rank)
        res = run(h, part, scene_ids=[s for s in scene_ids if s in mine], device=device, warning_file=warn)
        man = shard.gather_manifest({"rank": rank, "part": part, "scenes": len(mine)})
        if rank == 0:
            parts = [m["part"] for m in man]
            if a.table == "object_visibility":
                merged = {}
                for p in parts:
                    with open(p, "rb") as f:
                        merged.update(pickle.load(f))
                with open(out, "wb") as f:
                    pickle.dump({s: merged[s] for s in scene_ids if s in merged}, f)
                res = {"scenes": len(merged), "file": out}
            else:
                res = {"rows": merge_parquet_parts(parts, out, scene_ids), "file": out}
                if a.table == "relations":
                    nz = [p.replace(".parquet", "_nonzero.parquet") for p in parts]
                    res["rows_nonzero"] = merge_parquet_parts(nz, out.replace(".parquet", "_nonzero.parquet"), scene_ids)
        if rank == 0:                                        # the reference writes ONE warning file per split
            shard.merge_text_parts([shard.part_path(warn, r) for r in range(1, world)], warn)
        results[split] = res
        import torch.distributed as dist
        dist.barrier()
    return results


if __name__ == "__main__":
    main()
