"""CPU arm of bench.py: the UNMODIFIED reference's own functions annotating independent frame pairs.

TEST / MEASUREMENT INFRASTRUCTURE (only ``bench.py``'s ``--impl reference`` leg and its ``cpu_baseline`` sample call this).
The reference code is imported through ``oracle/ref_shim.py`` from ``/root/reference`` (build container) or from the
byte-for-byte copy ``oracle/build_ref.py`` staged under ``oracle/_ref`` (GPU box).  Nothing of this repo's product code is
on the timed path: the synthetic scene is written to disk in the reference's ScanNet layout (``oracle/ref_layout.py``)
before the clock starts and is then only touched through ``SceneInfoHandler``.

Per pair: both frames ``project_3d_point_to_image`` + ``check_point_visibility`` (IH:313-386, as looped by
CFR:152-157 / MVI:93-100), ``extract_yaw_pitch`` (CFR:86-100), ``calculate_camera_overlap`` (CFR:102-137) + the pair's
distance / angle differences (CFR:180-189), ``np.intersect1d`` of the two visible-point lists and one sampled common point
through ``get_point_2d_coordinates_in_image`` in both frames (VC-coor:303-344), per-object visible counts as
``compute_object_visibility.py:110-136`` forms them (set intersections of the visible list with each object's points).

Two variants:
  * ``as_shipped``  — the stock handler: every call decodes the JPEG (image shape) and the PNG (depth), and every
    per-point call reloads ``aligned_points.npy`` (IH:133-139, 153-157, 272-276);
  * ``io_free``     — a subclass that serves the depth maps, the image shape and the points from RAM (the arithmetic alone).
"""
from __future__ import annotations

import os
import shutil
import tempfile
import time

import numpy as np


def _make_io_free_handler(ns, info_path, root, scene, depth):
    """Subclass of the reference handler whose three I/O getters read RAM copies (same return types)."""
    base = ns.info_handler.SceneInfoHandler
    pts6 = np.load(os.path.join(root, "scannet_instance_data", scene.scene_id, "aligned_points.npy"))
    depth_by_id = {image_id: depth[f] for f, image_id in enumerate(scene.image_ids)}
    shape = tuple(scene.image_hw)

    class RamHandler(base):
        def get_image_shape(self, scene_id, image_id=None):      # IH:133-139
            return shape

        def get_depth_image(self, scene_id, image_id):          # IH:153-157
            return depth_by_id[self.convert_image_id_to_key(image_id) if not isinstance(image_id, str) else image_id]

        def get_scene_points_align(self, scene_id):              # IH:272-276
            return pts6

    return RamHandler(info_path, posed_images_root=os.path.join(root, "posed_images"),
                      instance_data_root=os.path.join(root, "scannet_instance_data"))


def annotate_pairs(ns, handler, scene, first_pair, n_pairs, rng):
    """``n_pairs`` independent pairs (frames 2k, 2k+1; k from ``first_pair``) of one scene through the reference's functions."""
    sid = scene.scene_id
    cfr = ns.cfr
    pts = handler.get_scene_points_align(sid)[:, :3]                                     # CFR:143-145
    inst = handler.get_scene_instance_mask(sid)
    obj_points = [set(np.flatnonzero(inst == o + 1).tolist()) for o in range(scene.num_objects)]   # IH:250-261
    done = 0
    for k in range(first_pair, first_pair + n_pairs):
        ids = (scene.image_ids[2 * k], scene.image_ids[2 * k + 1])
        in_bounds, yaw, pitch, pos = {}, {}, {}, {}
        for image_id in ids:                                                             # CFR:152-171
            E = handler.get_extrinsic_matrix_align(sid, image_id)
            p2d, depth = handler.project_3d_point_to_image(sid, image_id, pts)
            in_bounds[image_id] = handler.check_point_visibility(sid, image_id, p2d, depth)
            yaw[image_id], pitch[image_id] = cfr.extract_yaw_pitch(E)
            pos[image_id] = E[:3, 3]
        a, b = ids
        rel = {"overlap": cfr.calculate_camera_overlap(in_bounds, a, b),                 # CFR:176-189
               "yaw": yaw[b] - yaw[a], "pitch": pitch[b] - pitch[a], "distance": np.linalg.norm(pos[b] - pos[a])}
        va, vb = np.flatnonzero(in_bounds[a]).tolist(), np.flatnonzero(in_bounds[b]).tolist()   # MVI:101-103
        common = np.intersect1d(va, vb)                                                  # VC-coor:303
        if len(common):
            p = int(common[int(rng.integers(0, len(common)))])
            H, W = handler.get_image_shape(sid)
            for image_id in ids:                                                         # VC-coor:323-344
                xy = handler.get_point_2d_coordinates_in_image(sid, image_id, p)[0]
                _ = (round(xy[0] / W * 1000), round(xy[1] / H * 1000))
        for image_id, vis in ((a, va), (b, vb)):                                         # COV:110-136
            vs = set(vis)
            for o, op in enumerate(obj_points):
                cnt = len(vs & op)
                _ = cnt / max(len(op), 1) * 100
        done += 1
        del rel
    return done


def worker(job):
    """One process = one scene, like the reference's Pool-by-scene (make_visibility_info.py:151).
    job = (seed, n_points, n_pairs, variants) -> {variant: (pairs, seconds)}; scene set-up, file writes and one warm-up
    pair per variant are untimed (steady-state rate: the per-scene object-point sets of COV:110-118 ARE inside the clock)."""
    seed, n_points, n_pairs, variants = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from multi_spatialmllm_b200 import synth
    from oracle import geometry as og
    from oracle import ref_layout, ref_shim
    ns = ref_shim.import_reference()
    s = synth.make_scene(seed=seed, n_points=n_points, n_frames=2 * (n_pairs + 1), n_objects=12)   # + one warm-up pair
    pts = s.points.astype(np.float64)
    H, W = s.image_hw
    E = [og.aligned_pose(s.axis_align, s.c2w[f]) for f in range(s.n_frames)]
    # depth maps are INPUTS (the GPU arm renders them at set-up too): our z-buffer definition + 1 mm, before the clock starts
    s.depth = np.stack([og.depth_to_u16(og.zbuffer_render(pts, s.K, E[f], H, W, *s.depth_hw), 1) for f in range(s.n_frames)])
    root = tempfile.mkdtemp(prefix="msp_ref_arm_")
    out = {}
    try:
        info_path = ref_layout.write_split([s], root)
        for variant in variants:
            if variant == "as_shipped":
                h = ns.info_handler.SceneInfoHandler(info_path, posed_images_root=os.path.join(root, "posed_images"),
                                                     instance_data_root=os.path.join(root, "scannet_instance_data"))
            else:
                h = _make_io_free_handler(ns, info_path, root, s, s.depth)
            rng = np.random.default_rng(seed)
            annotate_pairs(ns, h, s, n_pairs, 1, rng)          # untimed warm-up pair (first-touch pages, lazy imports, BLAS init)
            t0 = time.perf_counter()
            n = annotate_pairs(ns, h, s, 0, n_pairs, rng)
            out[variant] = (n, time.perf_counter() - t0)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return out


def noop(_):
    import numpy  # noqa: F401
    from oracle import ref_shim
    ref_shim.import_reference()
    return 0


def run_sample(n_points: int, pairs_per_worker: int, workers: int, seed0: int = 0, variants=("io_free",)):
    """``workers`` processes x ``pairs_per_worker`` pairs; returns {variant: (pairs, slowest worker's seconds)} and wall."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    jobs = [(seed0 + w, n_points, pairs_per_worker, tuple(variants)) for w in range(workers)]
    with ctx.Pool(workers) as pool:
        pool.map(noop, range(workers))                 # interpreters + reference import before the clock
        t0 = time.perf_counter()
        res = pool.map(worker, jobs)
        wall = time.perf_counter() - t0
    out = {}
    for v in variants:
        out[v] = (sum(r[v][0] for r in res), max(r[v][1] for r in res))
    return out, wall
