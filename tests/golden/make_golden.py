"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference through oracle/ref_shim.py, writes seeded synthetic
scenes in the reference's disk layout (oracle/ref_layout.py) into a temp dir,
runs the reference's own hot-path functions on them and stores inputs+outputs:

  scannet_small.npz   per-frame xy / z / visibility (IH:313-386), MVI.process_scene
                      tables (MVI:75-125), CFR.process_scene pair records (CFR:139-197),
                      get_point_2d_coordinates_in_image spot checks (IH:291-305)
  tapvid_small.npz    OM world transform / project_point / pair distances, by importing
                      single_object_movement_engine_coord.TwoFrameVideoQAEngine

The reference ships no fixtures of its own — these files are the pin.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from multi_spatialmllm_b200 import synth  # noqa: E402
from oracle import geometry as og  # noqa: E402
from oracle import ref_layout, ref_shim  # noqa: E402


def render_depth(scene, rng, margin_mm=1, hole_frac=0.03, noise_mm=3):
    """Stand-in sensor depth: z-buffer of the scene's own points, then holes and mm noise."""
    H, W = scene.image_hw
    dh, dw = scene.depth_hw
    out = np.zeros((scene.n_frames, dh, dw), np.uint16)
    pts = scene.points.astype(np.float64)
    for f in range(scene.n_frames):
        if not og.pose_is_valid(scene.c2w[f]):
            continue
        E = og.aligned_pose(scene.axis_align, scene.c2w[f])
        zb = og.zbuffer_render(pts, scene.K, E, H, W, dh, dw)
        d = og.depth_to_u16(zb, margin_mm).astype(np.int64)
        filled = d > 0
        d = d + rng.integers(-noise_mm, noise_mm + 1, d.shape) * filled
        d[rng.random(d.shape) < hole_frac] = 0
        out[f] = np.clip(d, 0, 65535).astype(np.uint16)
    return out


def make_scannet(ref, tmp):
    rng = np.random.default_rng(7)
    scene = synth.make_scene(seed=11, n_points=6000, n_frames=9, n_objects=5,
                             scene_id="scene0011_00", invalid_pose_every=5)
    scene.depth = render_depth(scene, rng)
    info_path = ref_layout.write_split([scene], tmp)
    IH = ref.info_handler
    with contextlib.redirect_stdout(io.StringIO()):
        h = IH.SceneInfoHandler(info_path,
                                posed_images_root=os.path.join(tmp, "posed_images"),
                                instance_data_root=os.path.join(tmp, "scannet_instance_data"))
    sid = scene.scene_id
    valid_ids = h.get_all_extrinsic_valid_image_ids(sid)
    pts = h.get_scene_points_align(sid)[:, :3]
    xy, z, mask, E_al = [], [], [], []
    with np.errstate(all="ignore"):
        for image_id in valid_ids:
            a, b = h.project_3d_point_to_image(sid, image_id, pts)
            m = h.check_point_visibility(sid, image_id, a, b)
            xy.append(np.array(a)); z.append(np.array(b)); mask.append(m)
            E_al.append(h.get_extrinsic_matrix_align(sid, image_id))
    # L2 tables through the reference's own per-scene workers
    warn = os.path.join(tmp, "warn.txt")
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        _, vis_tab = ref.mvi.process_scene(sid, h, warn)
        _, pair_tab = ref.cfr.process_scene(sid, h, warn)
    pair_keys = list(pair_tab.keys())
    pair_vals = np.array([[pair_tab[k]["overlap"], pair_tab[k]["distance"], pair_tab[k]["yaw"], pair_tab[k]["pitch"]]
                          for k in pair_keys])
    # per-point spot checks through get_point_2d_coordinates_in_image
    spot = []
    for fi, image_id in enumerate(valid_ids):
        vis_pts = np.where(mask[fi])[0]
        pick = vis_pts[:: max(1, len(vis_pts) // 12)][:12]
        for p in pick:
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                a, b = h.get_point_2d_coordinates_in_image(sid, image_id, int(p), align=True,
                                                           check_visible=True, return_depth=True)
            assert len(a) == 1
            u, v, zz = float(a[0][0]), float(a[0][1]), float(b[0])
            spot.append([fi, p, u, v, zz,
                         round((u / 1296) * 1000), round((v / 968) * 1000), round(zz * 1000), int(u), int(v)])
    i2p = vis_tab["image_to_points"]
    p2i = vis_tab["point_to_images"]
    i2p_flat = np.concatenate([np.asarray(i2p[i], dtype=np.int64) for i in valid_ids])
    i2p_off = np.cumsum([0] + [len(i2p[i]) for i in valid_ids])
    p2i_cnt = np.array([len(p2i[p]) for p in range(scene.n_points)])
    np.savez_compressed(
        os.path.join(HERE, "scannet_small.npz"),
        points=scene.points, instance_ids=scene.instance_ids, num_objects=scene.num_objects,
        axis_align=scene.axis_align, c2w=scene.c2w, K=scene.K, depth=scene.depth,
        image_ids=np.array(scene.image_ids), valid_ids=np.array(valid_ids),
        image_hw=np.array(scene.image_hw), depth_hw=np.array(scene.depth_hw),
        ref_xy=np.stack(xy), ref_z=np.stack(z), ref_mask=np.stack(mask), ref_E_aligned=np.stack(E_al),
        ref_pair_ids=np.array(pair_keys), ref_pair_vals=pair_vals,
        ref_i2p_flat=i2p_flat, ref_i2p_off=i2p_off, ref_p2i_cnt=p2i_cnt,
        ref_spot=np.array(spot, dtype=np.float64),
    )
    print("scannet_small.npz:", len(valid_ids), "valid frames,", int(np.stack(mask).sum()), "visible point-frames,",
          len(pair_keys), "pairs,", len(spot), "spot checks")


def make_tapvid(ref, tmp):
    with contextlib.redirect_stdout(io.StringIO()):
        from spatial_engine.object_movement import single_object_movement_engine_coord as om
    vid = synth.make_video(seed=3, n_frames=24, n_tracks=40, n_groups=3)
    T, N, _ = vid.tracks_xyz.shape
    H, W = vid.image_hw
    # OM:446-451 restated by the reference itself is inline code; replicate its exact calls
    c2w = np.linalg.inv(vid.w2c)
    hom = np.concatenate([vid.tracks_xyz, np.ones((T, N, 1))], axis=2)
    world = np.einsum("nij,nkj->nki", c2w, hom)[..., :3]
    eng = om.TwoFrameVideoQAEngine.__new__(om.TwoFrameVideoQAEngine)
    eng.object_not_moving_threshold = 0.01
    eng.camera_not_moving_threshold = 0.01
    proj = np.full((T, N, 2), np.nan)
    with contextlib.redirect_stdout(io.StringIO()):
        for t in range(T):
            for n in range(N):
                # float64 input pins the reference environment's semantics (numpy 1.26.4, requirements/data_engine.yaml:162):
                # there np.float32 + 1e-8 promotes to float64, under numpy>=2 (NEP 50) it would stay float32.
                r = eng.project_point(vid.tracks_xyz[t, n].astype(np.float64), vid.fxfycxcy, H, W)
                if r is not None:
                    proj[t, n] = r
    np.savez_compressed(
        os.path.join(HERE, "tapvid_small.npz"),
        tracks_xyz=vid.tracks_xyz, visibility=vid.visibility, w2c=vid.w2c, fxfycxcy=vid.fxfycxcy,
        image_hw=np.array(vid.image_hw), ref_world=world, ref_proj=proj,
    )
    print("tapvid_small.npz:", T, "frames", N, "tracks,", int(np.isfinite(proj[..., 0]).sum()), "valid projections")


def main():
    ref = ref_shim.import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        make_scannet(ref, tmp)
        make_tapvid(ref, tmp)


if __name__ == "__main__":
    main()
